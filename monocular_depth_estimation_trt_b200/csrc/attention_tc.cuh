// Fused softmax(Q K^T * scale) V on tcgen05 tensor cores with TMEM accumulators, head dim 64.
//
// One CTA = 128 query rows of one (image, head); it walks the keys in tiles of 128.  Two CTAs are resident per
// SM (256 TMEM columns, ~113 KB of shared memory and 256 threads each), so one CTA's exponentials overlap the
// other's TMEM traffic, MMAs and prologue without any explicit hand-shake between them.
//   warps 0-3   softmax group (thread = query row = TMEM lane)
//   warp 4      TMA producer: the Q tile once, then K and V tiles through a 3-stage ring
//   warp 5      MMA issuer (one thread):  S = Q K_j^T   (M128 N128 K64, both operands K-major in smem)
//                                         O += P V_j    (M128 N64 K128, P read from TMEM, V MN-major in smem)
//
// A softmax thread pulls its whole score row (128 fp32) out of TMEM in one go and hands the S buffer
// straight back, so the tensor core computes S of the next key tile while this one is exponentiated.
// O accumulates in TMEM across key tiles.  The running maximum is applied lazily: the row keeps a stale
// reference maximum and only when the true maximum has grown by more than 2^8 is O (and the running sum)
// rescaled -- a TMEM load / multiply / store by the same thread, before it releases P for the next MMA.
// Probabilities therefore stay <= 256, well inside the range of bf16 and fp16.  P never touches shared
// memory: the softmax thread writes its packed 16-bit row back to TMEM (tcgen05.st) and the P V MMA takes
// its A operand from there, which halves the shared-memory traffic of the kernel (the other resource that
// would saturate next to the SFUs).
// At head dim 64 the kernel is bound by the 16 ex2/clk/SM special-function rate, not by the tensor pipe.
#pragma once
#include <cuda/std/type_traits>

#include "attention_mma.cuh"   // AttnParams, fast_exp2
#include "ptx.cuh"

namespace mde {

constexpr int kAtcThreads = 256;   // softmax warpgroup + a warpgroup holding the TMA and MMA warps
constexpr int kAtcQBytes = 128 * 64 * 2;          // one 128 x 64 16-bit tile
constexpr int kAtcStages = 3;
constexpr int kAtcTmemCols = 256;  // S [0,128)  O [128,192)  P [192,256)
// smem: Q | K[stages] | V[stages] | barriers     (two CTAs per SM)
constexpr int kAtcSmemBytes = kAtcQBytes + 2 * kAtcStages * kAtcQBytes + 256;   // 114 944 B: two fit in 228 KB
constexpr float kAtcRescaleThreshold = 8.0f;      // log2 units

// Operand tile with the N (or M) index contiguous: rows of 128 bytes are K indices, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(16384 >> 4) << 16;   // LBO: next 64-wide MN block (unused for N = 64)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// 2^x for two values on the FMA pipe (Cody-Waite: x = n + f, |f| <= 0.5, 2^f by a minimax polynomial, 2^n added
// into the exponent field).  Relative error 1.0e-4 (degree 3, bf16 probabilities round at 2^-9) or 3.7e-6
// (degree 4, fp16).  The special-function unit delivers 16 ex2 per clock and SM and is what bounds this kernel
// at head dim 64; kPoly of every 8 element pairs take this path instead and run beside it.
template <int kDegree>
__device__ __forceinline__ void exp2_fma2(f32x2 x, float& p0, float& p1) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  x = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const f32x2 r = f2_add(x, f2_splat(12582912.0f));            // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const f32x2 n = f2_add(r, f2_splat(-12582912.0f));
  const f32x2 f = f2_fma(n, f2_splat(-1.0f), x);
  f32x2 q;
  if (kDegree == 3) {
    q = f2_fma(f, f2_splat(0.05592203512787819f), f2_splat(0.24264007806777954f));
    q = f2_fma(q, f, f2_splat(0.6931210160255432f));
    q = f2_fma(q, f, f2_splat(0.9999244809150696f));
  } else {
    q = f2_fma(f, f2_splat(0.009676037356257439f), f2_splat(0.05592203512787819f));
    q = f2_fma(q, f, f2_splat(0.2402210682630539f));
    q = f2_fma(q, f, f2_splat(0.6931210160255432f));
    q = f2_fma(q, f, f2_splat(1.0000001192092896f));
  }
  float q0, q1, r0, r1;
  f2_unpack(q, q0, q1);
  f2_unpack(r, r0, r1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

// kTrace: the same kernel with clock64 stamps of every phase of the softmax warps written to p.trace (lane 0 of each warp,
// the first kAtcTraceCtas CTAs): [cta][warp][0] kernel entry, [1] after the prologue sync, then 5 per key tile -- S available,
// S in registers, exponentials done, previous P V done, P stored and announced -- and after the last tile: O available, stored.
constexpr int kAtcTraceCtas = 2048, kAtcTraceSlots = 64;
template <typename T, int kPoly, bool kTrace = false>
__global__ void __launch_bounds__(kAtcThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // images from the last one down: the QKV GEMM swept the rows front to back, its last ~100 MB are still in the L2; and the
  // projection GEMM behind this kernel starts at row 0, which is then what was written last
  const int head = blockIdx.y, img = gridDim.z - 1 - blockIdx.z;
  const int q0 = blockIdx.x * 128;
  const int nkv = (p.ntok + 127) / 128;
  const int last_chunks = (p.ntok - (nkv - 1) * 128 + 31) / 32;   // 32-key chunks of the last key tile that hold real keys (1..4)
  const int row_base = img * p.ntok_q;     // first query row of this image
  const int kv_base = img * p.ntok;        // first key/value row of this image
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  int trace_n = 0;
  auto stamp = [&]() {
    if (kTrace) {
      if (lane == 0 && warp < 4 && cta_lin < kAtcTraceCtas && trace_n < kAtcTraceSlots)
        p.trace[(static_cast<long long>(cta_lin) * 4 + warp) * kAtcTraceSlots + trace_n] = clock64();
      ++trace_n;
    }
  };
  stamp();
  if (kTrace && lane == 0 && warp < 4 && cta_lin < kAtcTraceCtas) {      // which SM: the last slot of the warp's record
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[(static_cast<long long>(cta_lin) * 4 + warp) * kAtcTraceSlots + kAtcTraceSlots - 1] = smid;
  }

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&map_qkv);
    prefetch_tmap(&map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kAtcStages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_ready, 128); mbar_init(o_full, 1);
    fence_mbar_init();
    // the first loads go out before the TMEM allocation, the CTA-wide sync and the register re-partition: their latency
    // (q|k|v was written by the previous kernel, mostly to HBM) is the longest item of the CTA's prologue
    griddep_wait();
    mbar_arrive_expect_tx(q_full, kAtcQBytes);
    tma_load_2d(sQ, &map_qkv, q_full, head * 64, row_base + q0);
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
  stamp();

  // Register re-partition per warpgroup: the single-thread roles need almost nothing, a softmax thread
  // holds a 128-wide score row.  2 CTAs x 256 threads start at 128 registers each.
  if (warp == 4) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      // Q and key tile 0 were requested in the prologue.  K runs one tile ahead of V: a K stage is free as soon as its
      // S = Q K^T has been computed (early), a V stage only when its P V has (a whole softmax later); waiting for the V stage
      // first would hold the next K back and S of the next key tile with it.
      for (int j = 1; j <= nkv; ++j) {
        if (j < nkv) {
          const int st = j % kAtcStages;
          mbar_wait(&k_empty[st], ((j / kAtcStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], kAtcQBytes);
          tma_load_2d(sK + st * kAtcQBytes, &map_kv, &k_full[st], p.k_col0 + head * 64, kv_base + j * 128);
        }
        if (j >= 2) {
          const int i = j - 1, st = i % kAtcStages;
          mbar_wait(&v_empty[st], ((i / kAtcStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[st], kAtcQBytes);
          tma_load_2d(sV + st * kAtcQBytes, &map_kv, &v_full[st], p.v_col0 + head * 64, kv_base + i * 128);
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================== MMA issuer
    // The whole warp runs the loop (every lane waits on the barriers) and one elected lane issues.  Descriptors and addresses
    // are computed from warp-uniform values, so the compiler keeps them in uniform registers, where tcgen05.mma wants them, and
    // issues the MMAs of a product back to back; under `if (lane == 0)` it wrapped every MMA in an elect loop with register ->
    // uniform-register moves, ~110 clk per MMA (profiles/r02_attention_q3_traces.txt).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    {
      constexpr uint32_t idesc_o = umma_idesc_f16(Tr::kFmt, 128, 64) | (1u << 16);   // B (= V) is MN-major
      const uint32_t smem0 = smem_u32(atc_smem);
      const uint64_t qa = umma_desc_k_sw128(smem0);
      // the last key tile only spans the 32-key chunks that hold real keys: fewer S columns, fewer P V steps
      auto issue_s = [&](int st, int j) {
        const uint32_t idesc_s = umma_idesc_f16(Tr::kFmt, 128, j == nkv - 1 ? last_chunks * 32 : 128);
        const uint64_t b = umma_desc_k_sw128(smem0 + kAtcQBytes + st * kAtcQBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, qa + 2 * k, b + 2 * k, idesc_s, k != 0);
          tc_commit(s_full);
          tc_commit(&k_empty[st]);
        }
        __syncwarp();
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
        const uint64_t vb = umma_desc_mn_sw128(smem0 + (1 + kAtcStages + st) * kAtcQBytes);
        const int ksteps = j == nkv - 1 ? 2 * last_chunks : 8;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)   // 16 keys per step: 8 packed P columns, two 8-row groups of V (2048 bytes)
            if (k < ksteps) tc_mma_f16_ts(tmem_base + 128, tmem_base + 192 + 8 * k, vb + 128 * k, idesc_o, (j | k) != 0);
          tc_commit(o_full);
          tc_commit(&v_empty[st]);
        }
        __syncwarp();
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
    const float sl = p.scale_log2;

    // nch_tag: 32-key chunks of this tile that hold real keys (4 with every key valid = the fast path; the last key
    // tile of a row may have fewer and a ragged end).  A compile-time count keeps every register array statically indexed.
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
      mbar_arrive(s_free);                       // the tensor core may overwrite S now
      stamp();
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
      stamp();
      // ---- the previous product has read P (and, for a rescale, written O): only now may either change
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      stamp();
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
      stamp();
    };

    using cuda::std::integral_constant;
    const int n_full = p.ntok / 128;             // key tiles without a ragged end: no per-tile test inside the hot loop
    for (int j = 0; j < n_full; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      stamp();
      tile(integral_constant<int, 4>{}, cuda::std::true_type{}, j);
    }
    if (n_full < nkv) {
      const int j = n_full;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      stamp();
      if (last_chunks == 1) tile(integral_constant<int, 1>{}, cuda::std::false_type{}, j);
      else if (last_chunks == 2) tile(integral_constant<int, 2>{}, cuda::std::false_type{}, j);
      else if (last_chunks == 3) tile(integral_constant<int, 3>{}, cuda::std::false_type{}, j);
      else tile(integral_constant<int, 4>{}, cuda::std::false_type{}, j);
    }
    // ---- normalise and store this row (128 contiguous bytes)
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    stamp();
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
    stamp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAtcTmemCols);
  }
}

}  // namespace mde
