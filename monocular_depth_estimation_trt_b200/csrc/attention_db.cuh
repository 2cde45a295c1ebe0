// Fused softmax(Q K^T * scale) V on tcgen05, head dim 64 -- second generation: S DOUBLE-BUFFERED in TMEM.
//
// Same roles and the same two-CTAs-per-SM residency as attention_tc.cuh (warps 0-3 softmax, thread = query row = TMEM lane;
// warp 4 TMA producer; warp 5 MMA issuer), but the 256 TMEM columns of a CTA are cut differently:
//
//     S0 [0, 96)   S1 [96, 192)   O [192, 256)          key tiles of 96 (three 32-key chunks)
//
// and P(j) is written over the first 48 columns of the S buffer it was computed from (a softmax thread owns its lane: once
// its score row is in registers nobody else needs those columns).  What that buys, per the phase trace of the first kernel
// (profiles/r01_attention_phase_trace.txt):
//   * S(j+1) is in TMEM before the softmax threads have finished tile j -- it was issued right after P V(j-1), into the
//     other buffer -- so a warp that runs ahead of its siblings no longer waits for the slowest one to free the single S
//     buffer (9.5 % of a CTA's life);
//   * storing P(j) does not wait for P V(j-1) any more (it used to share one P buffer: 4.9 %); only the rare O rescale does;
//   * kSpec: the exponentials of tile j >= 1 are taken against the (stale, lazily updated) reference maximum straight away,
//     chunk by chunk while the next chunk's TMEM load is in flight, and the row maximum is tracked beside them instead of in
//     a pass of its own in front of them.  If it turns out that the maximum grew by more than 2^8 (the same lazy-rescale
//     threshold as before) the tile is redone on the exact path from the S buffer, which is still intact.
// The MMA thread issues S(0), S(1), then per key tile: P V(j), S(j+2).  tcgen05.mma instructions of one thread execute in
// issue order, so S(j+2) overwrites buffer j & 1 only after P V(j) has read P(j) from it.
#pragma once
#include <cuda/std/type_traits>

#include "attention_tc.cuh"   // exp2_fma2, umma_desc_mn_sw128, AttnParams

namespace mde {

constexpr int kAdbKeys = 96;                       // keys per tile
constexpr int kAdbChunks = kAdbKeys / 32;
constexpr int kAdbKvBytes = kAdbKeys * 64 * 2;     // one 96 x 64 16-bit tile
constexpr int kAdbStages = 4;
constexpr int kAdbTmemCols = 256;
constexpr int kAdbOCol = 192;
constexpr int kAdbSmemBytes = kAtcQBytes + 2 * kAdbStages * kAdbKvBytes + 256;   // 114 944 B: two CTAs per SM

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <typename T, int kPoly, bool kSpec>
__global__ void __launch_bounds__(kAtcThreads, 2)
attention_db_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(1024) uint8_t adb_smem[];
  if ((smem_u32(adb_smem) & 1023u) != 0) __trap();
  uint8_t* sQ = adb_smem;
  uint8_t* sK = sQ + kAtcQBytes;
  uint8_t* sV = sK + kAdbStages * kAdbKvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAdbStages * kAdbKvBytes);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = bars + 1;               // [stages]
  uint64_t* k_empty = k_full + kAdbStages;
  uint64_t* v_full = k_empty + kAdbStages;
  uint64_t* v_empty = v_full + kAdbStages;
  uint64_t* s_full = v_empty + kAdbStages;   // [2] S(j) ready in buffer j & 1 (tcgen05.commit)
  uint64_t* p_ready = s_full + 2;            // [2] P(j) stored over S(j), O rescaled if needed (128 arrivals)
  // O += P V_j complete (tcgen05.commit) on o_full[j & 1].  Two barriers because the softmax threads only look at them when they
  // rescale O and at the very end: by then a single barrier could be TWO phases behind the one they wait for, and a parity
  // wait cannot tell phase j - 2 from phase j.  With one barrier per tile parity the lag is at most one phase of that barrier
  // (S(j) in TMEM implies P V(j-2) complete).
  uint64_t* o_full = p_ready + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, img = blockIdx.z;
  const int q0 = blockIdx.x * 128;
  const int nkv = (p.ntok + kAdbKeys - 1) / kAdbKeys;
  const int last_chunks = (p.ntok - (nkv - 1) * kAdbKeys + 31) / 32;   // live 32-key chunks of the last key tile (1..3)
  const int row_base = img * p.ntok_q;
  const int kv_base = img * p.ntok;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAdbStages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); }
    mbar_init(&o_full[0], 1); mbar_init(&o_full[1], 1);
    fence_mbar_init();
    // first loads before the TMEM allocation and the CTA-wide sync: their latency is the longest item of the prologue
    griddep_wait();
    mbar_arrive_expect_tx(q_full, kAtcQBytes);
    tma_load_2d(sQ, &map_q, q_full, head * 64, row_base + q0);
    mbar_arrive_expect_tx(&k_full[0], kAdbKvBytes);
    tma_load_2d(sK, &map_kv, &k_full[0], p.k_col0 + head * 64, kv_base);
    if (nkv > 1) {
      mbar_arrive_expect_tx(&k_full[1], kAdbKvBytes);
      tma_load_2d(sK + kAdbKvBytes, &map_kv, &k_full[1], p.k_col0 + head * 64, kv_base + kAdbKeys);
    }
    mbar_arrive_expect_tx(&v_full[0], kAdbKvBytes);
    tma_load_2d(sV, &map_kv, &v_full[0], p.v_col0 + head * 64, kv_base);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kAdbTmemCols);
    tmem_relinquish();
  }
  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 4) {
    // ===================================================== TMA producer: K runs two tiles ahead of V
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      for (int j = 2; j <= nkv; ++j) {
        if (j < nkv) {
          const int st = j % kAdbStages;
          mbar_wait(&k_empty[st], ((j / kAdbStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], kAdbKvBytes);
          tma_load_2d(sK + st * kAdbKvBytes, &map_kv, &k_full[st], p.k_col0 + head * 64, kv_base + j * kAdbKeys);
        }
        const int i = j - 1, st = i % kAdbStages;
        mbar_wait(&v_empty[st], ((i / kAdbStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kAdbKvBytes);
        tma_load_2d(sV + st * kAdbKvBytes, &map_kv, &v_full[st], p.v_col0 + head * 64, kv_base + i * kAdbKeys);
      }
    }
  } else if (warp == 5) {
    // ===================================================== MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      constexpr uint32_t idesc_o = umma_idesc_f16(Tr::kFmt, 128, 64) | (1u << 16);   // B (= V) is MN-major
      auto issue_s = [&](int j) {
        const int st = j % kAdbStages;
        mbar_wait(&k_full[st], (j / kAdbStages) & 1);
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_f16(Tr::kFmt, 128, j == nkv - 1 ? last_chunks * 32 : kAdbKeys);
        const uint64_t a = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t b = umma_desc_k_sw128(smem_u32(sK + st * kAdbKvBytes));
        const uint32_t d = tmem_base + (j & 1) * kAdbKeys;
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(d, a + 2 * k, b + 2 * k, idesc_s, k != 0);
        tc_commit(&s_full[j & 1]);
        tc_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (nkv > 1) issue_s(1);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kAdbStages;
        mbar_wait(&v_full[st], (j / kAdbStages) & 1);
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);      // P(j) over S(j), O rescaled
        tc_fence_after();
        const uint64_t vb = umma_desc_mn_sw128(smem_u32(sV + st * kAdbKvBytes));
        const uint32_t p_tmem = tmem_base + (j & 1) * kAdbKeys;
        const int ksteps = j == nkv - 1 ? 2 * last_chunks : 2 * kAdbChunks;
#pragma unroll
        for (int k = 0; k < 2 * kAdbChunks; ++k)   // 16 keys per step: 8 packed P columns, two 8-row groups of V (2048 bytes)
          if (k < ksteps) tc_mma_f16_ts(tmem_base + kAdbOCol, p_tmem + 8 * k, vb + 128 * k, idesc_o, (j | k) != 0);
        tc_commit(&o_full[j & 1]);
        tc_commit(&v_empty[st]);
        if (j + 2 < nkv) issue_s(j + 2);           // into the buffer P V(j) has just read: ordered behind it
      }
    }
  } else if (warp >= 6) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else {
    // ===================================================== softmax group (thread = query row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int r = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t o_addr = tmem_base + lane_base + kAdbOCol;
    float m_ref = -INFINITY;      // (possibly stale) maximum the probabilities are taken against
    float l_run = 0.f;
    const float sl = p.scale_log2;

    // p0, p1 = exp2 of a packed pair; pair index decides between the SFU and the FMA-pipe polynomial
    auto exp_pair = [&](f32x2 xs, int pair, float& p0, float& p1) {
      if ((pair & 7) < kPoly) {
        exp2_fma2<Tr::kFmt == 1 ? 3 : 4>(xs, p0, p1);
      } else {
        float x0, x1;
        f2_unpack(xs, x0, x1);
        p0 = fast_exp2(x0);
        p1 = fast_exp2(x1);
      }
    };
    auto publish = [&](uint32_t s_addr, const uint32_t (&pk0)[32], const uint32_t (&pk1)[16], int nch, const f32x2 (&rs2)[4], int b) {
      tmem_st_32x32b_x32(s_addr, pk0);
      if (nch > 2) tmem_st_32x32b_x16(s_addr + 32, pk1);
      tmem_st_wait();
      float a0, a1, b0, b1;
      f2_unpack(f2_add(rs2[0], rs2[1]), a0, a1);
      f2_unpack(f2_add(rs2[2], rs2[3]), b0, b1);
      l_run += (a0 + a1) + (b0 + b1);
      tc_fence_before();            // TMEM stores (P, rescaled O) are ordered before the MMA that reads them
      mbar_arrive(&p_ready[b]);
    };

    // ---- exact path: whole score row in registers, row maximum first, lazy rescale of O
    auto tile_exact = [&](auto nch_tag, auto full_tag, int j) {
      constexpr bool kFull = decltype(full_tag)::value;
      constexpr int nch = decltype(nch_tag)::value;
      const int b = j & 1;
      const uint32_t s_addr = tmem_base + lane_base + b * kAdbKeys;
      const int nvalid = kFull ? kAdbKeys : p.ntok - j * kAdbKeys;
      uint32_t raw[kAdbChunks][32];
#pragma unroll
      for (int ch = 0; ch < kAdbChunks; ++ch)
        if (ch < nch) tmem_ld_32x32b_x32(s_addr + ch * 32, raw[ch]);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < kAdbChunks; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch < nch && (kFull || ch < nch - 1 || ch * 32 + i < nvalid)) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[ch][i]));
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const bool grow = (mx - m_ref) * sl > kAtcRescaleThreshold;      // always on the first tile
      const float msl_new = (grow ? mx : m_ref) * sl;
      f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk0[32], pk1[16];
#pragma unroll
      for (int ch = 0; ch < kAdbChunks; ++ch) {
        if (ch < nch) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const f32x2 xs = f2_fma(f2_pack(__uint_as_float(raw[ch][i]), __uint_as_float(raw[ch][i + 1])), f2_splat(sl), f2_splat(-msl_new));
            float p0, p1;
            exp_pair(xs, i >> 1, p0, p1);
            if (!kFull && ch == nch - 1) {
              if (ch * 32 + i >= nvalid) p0 = 0.f;
              if (ch * 32 + i + 1 >= nvalid) p1 = 0.f;
            }
            rs2[(i >> 1) & 3] = f2_add(rs2[(i >> 1) & 3], f2_pack(p0, p1));
            if (ch < 2) pk0[ch * 16 + (i >> 1)] = Tr::pack2(p0, p1);
            else pk1[i >> 1] = Tr::pack2(p0, p1);
          }
        }
      }
      if (nch < 2) {
#pragma unroll
        for (int i = 16; i < 32; ++i) pk0[i] = 0u;       // stored with the live chunk, never multiplied (fewer P V steps)
      }
      if (__any_sync(0xffffffffu, grow)) {
        const float factor = grow ? fast_exp2((m_ref - mx) * sl) : 1.0f;
        if (grow) { m_ref = mx; l_run *= factor; }
        if (j > 0) {
          mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);   // the previous product has finished accumulating into O
          tc_fence_after();
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
      publish(s_addr, pk0, pk1, nch, rs2, b);
    };

    // ---- speculative path (full tiles, j >= 1): exponentials against the stale maximum while the chunks stream in
    auto tile_spec = [&](int j) -> bool {
      const int b = j & 1;
      const uint32_t s_addr = tmem_base + lane_base + b * kAdbKeys;
      const float msl = m_ref * sl;
      f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk0[32], pk1[16];
      float x4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t raw[2][32];
      tmem_ld_32x32b_x32(s_addr, raw[0]);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < kAdbChunks; ++ch) {
        if (ch + 1 < kAdbChunks) tmem_ld_32x32b_x32(s_addr + (ch + 1) * 32, raw[(ch + 1) & 1]);   // in flight during the arithmetic
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const f32x2 xs = f2_fma(f2_pack(__uint_as_float(raw[ch & 1][i]), __uint_as_float(raw[ch & 1][i + 1])), f2_splat(sl), f2_splat(-msl));
          float x0, x1, p0, p1;
          f2_unpack(xs, x0, x1);
          x4[(i >> 1) & 3] = fmaxf(fmaxf(x4[(i >> 1) & 3], x0), x1);
          exp_pair(xs, i >> 1, p0, p1);
          rs2[(i >> 1) & 3] = f2_add(rs2[(i >> 1) & 3], f2_pack(p0, p1));
          if (ch < 2) pk0[ch * 16 + (i >> 1)] = Tr::pack2(p0, p1);
          else pk1[i >> 1] = Tr::pack2(p0, p1);
        }
        if (ch + 1 < kAdbChunks) tmem_ld_wait();
      }
      const float xmax = fmaxf(fmaxf(x4[0], x4[1]), fmaxf(x4[2], x4[3]));
      if (__any_sync(0xffffffffu, xmax > kAtcRescaleThreshold)) return false;      // S(j) is intact: redo on the exact path
      publish(s_addr, pk0, pk1, kAdbChunks, rs2, b);
      return true;
    };

    using cuda::std::integral_constant;
    const int n_full = p.ntok / kAdbKeys;
    for (int j = 0; j < n_full; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (kSpec && j > 0) {
        if (tile_spec(j)) continue;
      }
      tile_exact(integral_constant<int, kAdbChunks>{}, cuda::std::true_type{}, j);
    }
    if (n_full < nkv) {
      const int j = n_full;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (last_chunks == 1) tile_exact(integral_constant<int, 1>{}, cuda::std::false_type{}, j);
      else if (last_chunks == 2) tile_exact(integral_constant<int, 2>{}, cuda::std::false_type{}, j);
      else tile_exact(integral_constant<int, 3>{}, cuda::std::false_type{}, j);
    }
    // ---- normalise and store this row (128 contiguous bytes)
    mbar_wait(&o_full[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
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
    tmem_dealloc(tmem_base, kAdbTmemCols);
  }
}

}  // namespace mde
