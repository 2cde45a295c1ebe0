// The default attention kernel (attention_tc.cuh) with the softmax of a 128-row query tile spread over EIGHT warps instead of
// four: warps w and w + 4 share the TMEM lane quarter w % 4 (the hardware lets both read it) and each takes 64 of the 128
// score columns of "its" rows.  Everything else -- one query tile per CTA, two CTAs per SM, 128-key tiles, S / O / P in
// TMEM, lazy rescale, 3-stage K/V ring -- is unchanged.  Why: with four softmax warps per CTA an SM sub-partition hosts two
// warps that run long dependent chains (FFMA2 -> MUFU.EX2 -> FADD2 / F2FP); ncu shows the issue slots 53 % busy and 1.4
// cycles of fixed-latency dependency stall per issued instruction.  Four warps per sub-partition, each with half the row,
// hide that latency; the price is one exchange of the row maximum per key tile between the two threads of a row
// (512 bytes of shared memory, a two-warp named barrier) and half-size TMEM transfers.
//   warps 0-7   softmax (thread = query row x column half)
//   warp 8      TMA producer        warp 9      MMA issuer        warps 10-11 idle (register donors)
#pragma once
#include <cuda/std/type_traits>

#include "attention_tc.cuh"
#include "attention_tc2q.cuh"   // named_bar_sync

namespace mde {

constexpr int kA8Threads = 384;
constexpr int kA8XchgBytes = 512;                  // bf16 row maxima of both halves / fp32 row sums of one half
constexpr int kA8SmemBytes = kAtcSmemBytes + kA8XchgBytes;   // 115 456 B: two CTAs still fit in 228 KB


template <typename T, int kPoly>
__global__ void __launch_bounds__(kA8Threads, 2)
attention_tc8w_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  uint16_t* xchg = reinterpret_cast<uint16_t*>(atc_smem + kAtcSmemBytes);   // [2 halves][128 rows] bf16 maxima (fp32 sums at the end)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, img = blockIdx.z;
  const int q0 = blockIdx.x * 128;
  const int nkv = (p.ntok + 127) / 128;
  const int last_chunks = (p.ntok - (nkv - 1) * 128 + 31) / 32;   // 32-key chunks of the last key tile that hold real keys (1..4)
  const int row_base = img * p.ntok_q;     // first query row of this image
  const int kv_base = img * p.ntok;        // first key/value row of this image

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&map_qkv);
    prefetch_tmap(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAtcStages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 256); mbar_init(p_ready, 256); mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) {
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
  if (warp == 8) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAtcQBytes);
      tma_load_2d(sQ, &map_qkv, q_full, head * 64, row_base + q0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kAtcStages;
        const uint32_t ph = (j / kAtcStages) & 1;
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
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
      for (int j = 0; j < nkv; ++j) {
        const int st = j % kAtcStages;
        if (j + 1 < nkv) {
          // S of the next key tile as soon as the softmax threads hold the current scores in registers
          const int st1 = (j + 1) % kAtcStages;
          mbar_wait(&k_full[st1], ((j + 1) / kAtcStages) & 1);
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          issue_s(st1, j + 1);
        }
        mbar_wait(&v_full[st], (j / kAtcStages) & 1);
        mbar_wait(p_ready, j & 1);                 // P(j) in TMEM, O rescaled
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
  } else if (warp >= 10) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");   // idle half of the producer warpgroup
  } else {
    // ===================================================== softmax group (thread = query row x column half)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + half * 64;
    const uint32_t o_addr = tmem_base + lane_base + 128 + half * 32;
    const uint32_t p_addr = tmem_base + lane_base + 192 + half * 32;
    const int bar_id = 1 + quarter;           // warps `quarter` and `quarter + 4` meet here (barrier 0 is __syncthreads)
    float m_ref = -INFINITY;      // (possibly stale) maximum the probabilities are taken against; identical in both halves
    float l_run = 0.f;            // this half's share of the row sum
    const float sl = p.scale_log2;

    // nch_tag: 32-key chunks of THIS half that hold real keys (2 on a full tile; 0..2 on the last, possibly ragged one)
    auto tile = [&](auto nch_tag, auto full_tag, int j) {
      constexpr bool kFull = decltype(full_tag)::value;
      constexpr int nch = decltype(nch_tag)::value;
      const int nvalid = kFull ? 64 : p.ntok - j * 128 - half * 64;     // valid keys among this half's 64 columns
      uint32_t raw[2][32];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
        if (ch < nch) tmem_ld_32x32b_x32(s_addr + ch * 32, raw[ch]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                       // the tensor core may overwrite S once all 256 threads hold their part
      // ---- row maximum of this half, four independent chains
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch < nch && (kFull || ch * 32 + i < nvalid)) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[ch][i]));
      const float mx_own = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      // ---- exchange with the thread that holds the other half of this row.  The maximum only has to be the SAME number in
      // both threads and not below the true one: it travels as bf16 rounded up (the lazy rescale tolerates 2^8 of slack).
      const __nv_bfloat16 up = __float2bfloat16_ru(mx_own);
      xchg[half * 128 + r] = __bfloat16_as_ushort(up);
      named_bar_sync(bar_id, 64);
      const float mx = fmaxf(__bfloat162float(up), __bfloat162float(__ushort_as_bfloat16(xchg[(half ^ 1) * 128 + r])));
      named_bar_sync(bar_id, 64);                // both have read: the slots may be overwritten in the next tile
      // ---- lazy rescale: only when the maximum grew by more than 2^8 (always on the first tile)
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
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[ch * 16 + i] = 0u;
        }
      }
      // ---- the previous product has read P (and, for a rescale, written O): only now may either change
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, grow)) {
        const float factor = grow ? fast_exp2((m_ref - mx) * sl) : 1.0f;
        if (grow) { m_ref = mx; l_run *= factor; }
        if (j > 0) {                             // each half rescales its 32 of O's 64 columns
          uint32_t o[32];
          tmem_ld_32x32b_x32(o_addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
          tmem_st_32x32b_x32(o_addr, o);
        }
      }
      if (nch > 0) tmem_st_32x32b_x32(p_addr, pk);
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

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      using cuda::std::integral_constant;
      if (j * 128 + 128 <= p.ntok) {
        tile(integral_constant<int, 2>{}, cuda::std::true_type{}, j);
      } else {
        const int mine = min(max(last_chunks - 2 * half, 0), 2);        // this half's live 32-key chunks of the last tile
        if (mine == 2) tile(integral_constant<int, 2>{}, cuda::std::false_type{}, j);
        else if (mine == 1) tile(integral_constant<int, 1>{}, cuda::std::false_type{}, j);
        else tile(integral_constant<int, 0>{}, cuda::std::false_type{}, j);
      }
    }
    // ---- total row sum = both halves' shares (fp32 through the same 512 bytes, one direction at a time)
    float* xf = reinterpret_cast<float*>(xchg);
    if (half == 1) xf[r] = l_run;
    named_bar_sync(bar_id, 64);
    float other = 0.f;
    if (half == 0) { other = xf[r]; }
    named_bar_sync(bar_id, 64);
    if (half == 0) xf[r] = l_run;
    named_bar_sync(bar_id, 64);
    if (half == 1) other = xf[r];
    // ---- normalise and store this thread's 32 of the row's 64 output features (64 contiguous bytes)
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / (l_run + other);
    const int n = q0 + r;
    T* gout = static_cast<T*>(p.out) + (static_cast<long long>(row_base) + n) * p.D + head * 64 + half * 32;
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_addr, o);
      tmem_ld_wait();
      if (n < p.ntok_q) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = Tr::pack2(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
          u.y = Tr::pack2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
          u.z = Tr::pack2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
          u.w = Tr::pack2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(gout + c * 8) = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAtcTmemCols);
  }
}

}  // namespace mde
