// The default attention kernel (attention_tc.cuh) working through kItems query tiles of one (image, head) per CTA, one after
// the other.  A CTA's prologue -- launch, barrier set-up, TMEM allocation, CTA sync, register re-partition and above all the
// HBM latency of Q and the first K / V tile before the first S = Q K^T exists -- costs the softmax warps about a tenth of a
// one-tile CTA's life (profiles/README.md, "Attention, second pass").  Here only the first item pays it: the producer replaces
// Q as soon as the item's last S has been computed (tcgen05.commit on `q_free`, about one key tile before the softmax threads
// finish), the K/V ring simply keeps streaming (the same head's keys again: L2 hits), and the MMA thread issues the next
// item's first S while the current item's last tile is still in the exponentials.  The only new dependency is `o_free`:
// the next item's first P V overwrites the O accumulator, so it waits until the softmax threads have read O out of TMEM.
#pragma once
#include "attention_tc.cuh"

namespace mde {

template <typename T, int kPoly, int kItems>
__global__ void __launch_bounds__(kAtcThreads, 2)
attention_tcq_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(1024) uint8_t atc_smem[];   // 128-byte-swizzled operand tiles need 1024-byte alignment
  if ((smem_u32(atc_smem) & 1023u) != 0) __trap();
  uint8_t* sQ = atc_smem;
  uint8_t* sK = sQ + kAtcQBytes;
  uint8_t* sV = sK + kAtcStages * kAtcQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAtcStages * kAtcQBytes);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [stages]
  uint64_t* k_empty = k_full + kAtcStages;
  uint64_t* v_full = k_empty + kAtcStages;
  uint64_t* v_empty = v_full + kAtcStages;
  uint64_t* s_full = v_empty + kAtcStages; // S ready in TMEM (tcgen05.commit)
  uint64_t* s_free = s_full + 1;           // S copied to registers (128 arrivals)
  uint64_t* p_ready = s_free + 1;          // P in TMEM, O rescaled if needed (128 arrivals)
  uint64_t* o_full = p_ready + 1;          // O += P V_j complete (tcgen05.commit)
  uint64_t* q_free = o_full + 1;           // every S = Q K^T of the item has completed: the Q tile may be replaced (tcgen05.commit)
  uint64_t* o_free = q_free + 1;           // the item's O has been read out of TMEM (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, img = blockIdx.z;
  const int q_tiles = (p.ntok_q + 127) / 128;
  const int first_tile = blockIdx.x * kItems;
  const int n_items = min(kItems, q_tiles - first_tile);     // query tiles this CTA works through, one after the other
  const int nkv = (p.ntok + 127) / 128;
  const int G = n_items * nkv;               // key tiles over all items: barrier phases count these
  const int last_chunks = (p.ntok - (nkv - 1) * 128 + 31) / 32;   // 32-key chunks of the last key tile that hold real keys (1..4)
  const int row_base = img * p.ntok_q;     // first query row of this image
  const int kv_base = img * p.ntok;        // first key/value row of this image

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&map_qkv);
    prefetch_tmap(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAtcStages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_ready, 128); mbar_init(o_full, 1);
    mbar_init(q_free, 1); mbar_init(o_free, 128);
    fence_mbar_init();
    // the first loads go out before the TMEM allocation, the CTA-wide sync and the register re-partition: their latency
    // (q|k|v was written by the previous kernel, mostly to HBM) is the longest item of the CTA's prologue
    griddep_wait();
    mbar_arrive_expect_tx(q_full, kAtcQBytes);
    tma_load_2d(sQ, &map_qkv, q_full, head * 64, row_base + first_tile * 128);
    mbar_arrive_expect_tx(&k_full[0], kAtcQBytes);
    tma_load_2d(sK, &map_kv, &k_full[0], p.k_col0 + head * 64, kv_base);
    mbar_arrive_expect_tx(&v_full[0], kAtcQBytes);
    tma_load_2d(sV, &map_kv, &v_full[0], p.v_col0 + head * 64, kv_base);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kAtcTmemCols);
    tmem_relinquish();
  }
  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                          // Q / K / V come from the previous kernel

  // Register re-partition per warpgroup: the single-thread roles need almost nothing, a softmax thread
  // holds a 128-wide score row.  2 CTAs x 256 threads start at 128 registers each.
  if (warp == 4) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      for (int g = 1; g < G; ++g) {                // Q of the first item and key tile 0 were requested in the prologue
        const int j = g % nkv, it = g / nkv;
        if (j == 0) {
          // next query tile: its Q replaces the current one as soon as the last S of the current item is done, about one
          // key tile before the softmax threads get there
          mbar_wait(q_free, (it - 1) & 1);
          mbar_arrive_expect_tx(q_full, kAtcQBytes);
          tma_load_2d(sQ, &map_qkv, q_full, head * 64, row_base + (first_tile + it) * 128);
        }
        const int st = g % kAtcStages;
        const uint32_t ph = (g / kAtcStages) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kAtcQBytes);
        tma_load_2d(sK + st * kAtcQBytes, &map_kv, &k_full[st], p.k_col0 + head * 64, kv_base + j * 128);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kAtcQBytes);
        tma_load_2d(sV + st * kAtcQBytes, &map_kv, &v_full[st], p.v_col0 + head * 64, kv_base + j * 128);
      }
    }
  } else if (warp == 5) {
    // ===================================================== MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      constexpr uint32_t idesc_o = umma_idesc_f16(Tr::kFmt, 128, 64) | (1u << 16);   // B (= V) is MN-major
      // the last key tile only spans the 32-key chunks that hold real keys: fewer S columns, fewer P V steps
      auto issue_s = [&](int st, int j) {
        const uint32_t idesc_s = umma_idesc_f16(Tr::kFmt, 128, j == nkv - 1 ? last_chunks * 32 : 128);
        const uint64_t a = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t b = umma_desc_k_sw128(smem_u32(sK + st * kAtcQBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, a + 2 * k, b + 2 * k, idesc_s, k != 0);
        tc_commit(s_full);
        tc_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      if (nkv == 1) tc_commit(q_free);
      for (int g = 0; g < G; ++g) {
        const int j = g % nkv, it = g / nkv;
        const int st = g % kAtcStages;
        if (g + 1 < G) {
          // S of the next key tile (of this item or of the next one) as soon as the softmax threads hold the current scores
          const int g1 = g + 1, j1 = g1 % nkv, st1 = g1 % kAtcStages;
          mbar_wait(&k_full[st1], (g1 / kAtcStages) & 1);
          if (j1 == 0) mbar_wait(q_full, (g1 / nkv) & 1);
          mbar_wait(s_free, g & 1);
          tc_fence_after();
          issue_s(st1, j1);
          if (j1 == nkv - 1) tc_commit(q_free);
        }
        mbar_wait(&v_full[st], (g / kAtcStages) & 1);
        mbar_wait(p_ready, g & 1);                 // P in TMEM, O rescaled
        if (j == 0 && it > 0) mbar_wait(o_free, (it - 1) & 1);   // the previous item's O has left TMEM
        tc_fence_after();
        const uint64_t vb = umma_desc_mn_sw128(smem_u32(sV + st * kAtcQBytes));
        const int ksteps = j == nkv - 1 ? 2 * last_chunks : 8;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 keys per step: 8 packed P columns, two 8-row groups of V (2048 bytes)
          if (k < ksteps) tc_mma_f16_ts(tmem_base + 128, tmem_base + 192 + 8 * k, vb + 128 * k, idesc_o, (j | k) != 0);
        tc_commit(o_full);
        tc_commit(&v_empty[st]);
      }
    }
  } else if (warp >= 6) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // idle half of the producer warpgroup
  } else {
    // ===================================================== softmax group (thread = query row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int r = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base;
    const uint32_t o_addr = tmem_base + lane_base + 128;
    const uint32_t p_addr = tmem_base + lane_base + 192;
    float m_ref = -INFINITY;      // (possibly stale) maximum the probabilities are taken against
    float l_run = 0.f;
    int gbase = 0;                // key tiles of the items already finished
    const float sl = p.scale_log2;

    // nch_tag: 32-key chunks of this tile that hold real keys (4 with every key valid = the fast path; the last key
    // tile of a row may have fewer and a ragged end).  A compile-time count keeps every register array statically indexed.
    auto tile = [&](auto nch_tag, auto full_tag, int j) {
      const int g = gbase + j;
      constexpr bool kFull = decltype(full_tag)::value;
      constexpr int nch = decltype(nch_tag)::value;
      const int nvalid = kFull ? 128 : p.ntok - j * 128;
      uint32_t raw[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch < nch) tmem_ld_32x32b_x32(s_addr + ch * 32, raw[ch]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                       // the tensor core may overwrite S now
      // ---- row maximum, four independent chains
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch < nch && (kFull || ch < nch - 1 || ch * 32 + i < nvalid)) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[ch][i]));   // only the last live chunk can be ragged
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      // ---- lazy rescale: only when the maximum grew by more than 2^8 (always on the first tile)
      const bool grow = (mx - m_ref) * sl > kAtcRescaleThreshold;
      // ---- P = exp2(S * sl - m * sl), packed to 16 bits in registers (the score registers die as we go)
      const float msl_new = (grow ? mx : m_ref) * sl;
      f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[2][32];
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
          if (!kFull && ch == nch - 1) {
            if (ch * 32 + i >= nvalid) p0 = 0.f;
            if (ch * 32 + i + 1 >= nvalid) p1 = 0.f;
          }
          rs2[(i >> 1) & 3] = f2_add(rs2[(i >> 1) & 3], f2_pack(p0, p1));
          pk[ch >> 1][(ch & 1) * 16 + (i >> 1)] = Tr::pack2(p0, p1);
        }
        }
      }
      // ---- the previous product has read P (and, for a rescale, written O): only now may either change
      if (j > 0) {
        mbar_wait(o_full, (g - 1) & 1);
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
      tc_fence_before();            // TMEM stores (P, rescaled O) are ordered before the MMA that reads them
      mbar_arrive(p_ready);
    };

    using cuda::std::integral_constant;
    const int n_full = p.ntok / 128;             // key tiles without a ragged end: no per-tile test inside the hot loop
    for (int it = 0; it < n_items; ++it) {
      m_ref = -INFINITY;
      l_run = 0.f;
      gbase = it * nkv;
      for (int j = 0; j < n_full; ++j) {
        mbar_wait(s_full, (gbase + j) & 1);
        tc_fence_after();
        tile(integral_constant<int, 4>{}, cuda::std::true_type{}, j);
      }
      if (n_full < nkv) {
        const int j = n_full;
        mbar_wait(s_full, (gbase + j) & 1);
        tc_fence_after();
        if (last_chunks == 1) tile(integral_constant<int, 1>{}, cuda::std::false_type{}, j);
        else if (last_chunks == 2) tile(integral_constant<int, 2>{}, cuda::std::false_type{}, j);
        else if (last_chunks == 3) tile(integral_constant<int, 3>{}, cuda::std::false_type{}, j);
        else tile(integral_constant<int, 4>{}, cuda::std::false_type{}, j);
      }
      // ---- normalise and store this row (128 contiguous bytes); O leaves TMEM first so that the next item's products can start
      mbar_wait(o_full, (gbase + nkv - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(o_addr, o0);
      tmem_ld_32x32b_x32(o_addr + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);
      const float inv = 1.0f / l_run;
      const int n = (first_tile + it) * 128 + r;
      T* gout = static_cast<T*>(p.out) + (static_cast<long long>(row_base) + n) * p.D + head * 64;
      if (n < p.ntok_q) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t* o = h ? o1 : o0;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAtcTmemCols);
  }
}

}  // namespace mde
