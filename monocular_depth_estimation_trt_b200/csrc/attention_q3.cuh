// Fused softmax(Q K^T * scale) V on tcgen05, head dim 64: THREE query tiles per CTA, one persistent CTA per SM.
//
// What bounds attention at head dim 64 is the special-function unit (16 ex2 per clock and SM).  A softmax warp cannot keep
// its sub-partition's SFU busy on its own: inside one warp the MUFU work and the packed-fp32 work add up (a lone warp needs
// ~1300 clk for the 768 MUFU clocks of a 128-key row, profiles/r02_attention_sm_overlap.txt), and between two rows the warp
// waits ~800 clk for TMEM traffic, the P V and the next S.  With two query tiles per SM (attention_tc.cuh) the SFU is busy
// 65 % of the time.  Here every sub-partition holds THREE softmax warps of three independent query tiles, so that (almost)
// always two of them are exponentiating.  All 512 TMEM columns belong to the one CTA:
//     query tile t:  S_t [160 t, 160 t + 96)   fp32 scores of a 96-key tile; P (16-bit, 48 columns) is stored over its own S
//                    O_t [160 t + 96, 160 t + 160)
//   warps 0-11  three softmax groups (thread = query row = TMEM lane)
//   warp 12     TMA producer: walks this CTA's work items, loads the Q tiles of an item and streams K / V
//               tiles (96 keys) through a 4-stage ring that the three query tiles SHARE (one third of the L2 reads)
//   warps 13-15 MMA issuers, one thread each: warp 13 + t owns query tile t and issues, per key tile, strictly
//               wait P_t(j);  O_t += P_t(j) V(j);  S_t(j + 1) = Q_t K(j + 1)^T      (S(j + 1) overwrites P(j), so it follows
//               P V(j) in issue order; the tensor pipe executes one thread's MMAs in order)
// A work item is (image, head, group of three consecutive query tiles); the last group of a row may hold fewer tiles.
// Items are handed out by a work counter when the op owns one (`counters` = {next item, CTAs done}, zero before the first
// launch; the last CTA through rearms it), else CTA b takes items b, b + gridDim.x, ...  The counter keeps the CTAs in one
// compact window of items however their speeds differ and is ~5 % faster (0.64 vs 0.67 ms at the batch-64 shape), but it is
// state that outlives the launch: two launches that may run side by side must not share one.  Engine contexts own a counter
// per attention op of their plan; the per-kernel entry points (mde_k_attention*), whose ops live for one call and may be
// frozen into the caller's CUDA graphs, take the static schedule.  Either way neighbouring CTAs work on neighbouring groups
// of the same (image, head): its K / V tiles are read from HBM once and from the L2 by the others.
#pragma once
#include <cuda/std/type_traits>

#include "attention_tc.cuh"   // exp2_fma2, umma_desc_mn_sw128, kAtcRescaleThreshold

namespace mde {

constexpr int kAq3Threads = 512;
constexpr int kAq3Keys = 96;
constexpr int kAq3QBytes = 128 * 64 * 2;
constexpr int kAq3KvBytes = kAq3Keys * 64 * 2;
constexpr int kAq3Stages = 4;
constexpr int kAq3TileCols = 160;
constexpr int kAq3OutBytes = 12 * 32 * 128;       // per softmax warp: 32 output rows of 128 bytes, staged for coalesced stores
constexpr int kAq3SmemBytes = 3 * kAq3QBytes + 2 * kAq3Stages * kAq3KvBytes + kAq3OutBytes + 512;   // 197 120 B

// kTrace: clock64 stamps into p.trace [CTA][16 rows][256 slots] (tools/attn_q3_trace.py): rows 0-11 the softmax warps (per key
// tile: S available, S in registers, exponentials done, P announced; per item: O available, O in registers, O staged, O stored), rows 12-14 the MMA
// threads' issue times of query tile 0-2 (S(0), PV(0), S(1), ...), row 15 the producer (per item: fetched, Q requested, K/V done).
constexpr int kAq3TraceRows = 16, kAq3TraceSlots = 256;
template <typename T, int kPoly, bool kTrace = false>
__global__ void __launch_bounds__(kAq3Threads, 1)
attention_q3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_out, const AttnParams p, unsigned int* __restrict__ counters) {
  using Tr = F16Traits<T>;
  extern __shared__ __align__(1024) uint8_t aq3_smem[];
  if ((smem_u32(aq3_smem) & 1023u) != 0) __trap();
  uint8_t* sQ = aq3_smem;
  uint8_t* sK = sQ + 3 * kAq3QBytes;
  uint8_t* sV = sK + kAq3Stages * kAq3KvBytes;
  uint8_t* sO = sV + kAq3Stages * kAq3KvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + kAq3OutBytes);
  uint64_t* q_full = bars;                        // [3]  Q tile of the item loaded
  uint64_t* q_empty = q_full + 3;                 // [3]  every S of the item computed: Q may be replaced
  uint64_t* k_full = q_empty + 3;                 // [stages]
  uint64_t* k_empty = k_full + kAq3Stages;
  uint64_t* v_full = k_empty + kAq3Stages;
  uint64_t* v_empty = v_full + kAq3Stages;
  uint64_t* s_full = v_empty + kAq3Stages;        // [3]  S_t in TMEM (and every earlier MMA complete, P V of the previous key tile included)
  uint64_t* p_ready = s_full + 3;                 // [3]  P_t in TMEM, O_t rescaled if needed (128 arrivals)
  uint64_t* o_full = p_ready + 3;                 // [3]  last P V of the item complete
  uint64_t* o_free = o_full + 3;                  // [3]  O_t copied to registers (128 arrivals)
  uint64_t* it_full = o_free + 3;                 // [4]  item ring
  uint64_t* it_empty = it_full + 4;               // [4]  15 arrivals: 12 softmax warps and the three MMA threads
  int* item_ring = reinterpret_cast<int*>(it_empty + 4);   // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(item_ring + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int trace_n[3] = {0, 0, 0};
  auto stamp = [&](int row, int ctr) {
    if (kTrace) {
      if (trace_n[ctr] < kAq3TraceSlots)
        p.trace[(static_cast<long long>(blockIdx.x) * kAq3TraceRows + row) * kAq3TraceSlots + trace_n[ctr]] = clock64();
      ++trace_n[ctr];
    }
  };
  const int nkv = (p.ntok + kAq3Keys - 1) / kAq3Keys;
  const int last_chunks = (p.ntok - (nkv - 1) * kAq3Keys + 31) / 32;   // 32-key chunks of the last key tile with real keys (1..3)
  const int q_tiles = (p.ntok_q + 127) / 128;
  const int n_groups = (q_tiles + 2) / 3;
  const int n_items = p.batch * p.heads * n_groups;
  // item -> (image, head, group); images from the last one down (see attention_tc.cuh).  The group index is skewed by the
  // (image, head) index: a row's last group may hold fewer query tiles than the others, and with gridDim.x a multiple of
  // n_groups (148 SMs, 4 groups at 1370 tokens) a CTA would otherwise meet the same group index in every round -- a quarter
  // of the CTAs all short items, the rest all long ones (measured: 0.70 instead of 0.64 ms).
  auto decode = [&](int item, int& img, int& head, int& g) {
    const int ih = item / n_groups;
    g = (item - ih * n_groups + ih) % n_groups;
    head = ih % p.heads;
    img = p.batch - 1 - ih / p.heads;
  };

  if (warp == 12 && lane == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_kv);
    prefetch_tmap(&map_out);
    for (int t = 0; t < 3; ++t) {
      mbar_init(&q_full[t], 1); mbar_init(&q_empty[t], 1); mbar_init(&s_full[t], 1); mbar_init(&p_ready[t], 128);
      mbar_init(&o_full[t], 1); mbar_init(&o_free[t], 128);
    }
    for (int i = 0; i < kAq3Stages; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 3); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 3);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&it_full[i], 1); mbar_init(&it_empty[i], 15); }
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                          // Q / K / V come from the previous kernel

  if (warp == 12) {
    // ===================================================== producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      int na0 = 0, na1 = 0, na2 = 0;       // items each query-tile slot has taken part in
      for (int n = 0;; ++n) {
        const int slot = n & 3;
        if (n >= 4) mbar_wait(&it_empty[slot], ((n >> 2) - 1) & 1);
        int item = counters ? static_cast<int>(atomicAdd(&counters[0], 1u)) : static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x);
        if (item >= n_items) item = -1;
        item_ring[slot] = item;
        mbar_arrive(&it_full[slot]);
        stamp(15, 0);
        if (item < 0) break;
        int img, head, g;
        decode(item, img, head, g);
        const int nq = min(3, q_tiles - 3 * g);
        const int row0 = img * p.ntok_q + g * 384, kv_base = img * p.ntok;
        // key tile 0 first: its ring stage has been free for a while, whereas the Q buffers are only released by the previous
        // item's last S -- the next item's first S needs both, and K must not queue up behind the wait for Q
        auto load_kv = [&](int j) {
          const int c = n * nkv + j, st = c % kAq3Stages, use = c / kAq3Stages;
          if (use > 0) mbar_wait(&k_empty[st], (use - 1) & 1);
          mbar_arrive_expect_tx(&k_full[st], kAq3KvBytes);
          tma_load_2d(sK + st * kAq3KvBytes, &map_kv, &k_full[st], p.k_col0 + head * 64, kv_base + j * kAq3Keys);
          if (use > 0) mbar_wait(&v_empty[st], (use - 1) & 1);
          mbar_arrive_expect_tx(&v_full[st], kAq3KvBytes);
          tma_load_2d(sV + st * kAq3KvBytes, &map_kv, &v_full[st], p.v_col0 + head * 64, kv_base + j * kAq3Keys);
        };
        load_kv(0);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (t < nq) {
            int& na = t == 0 ? na0 : t == 1 ? na1 : na2;
            if (na > 0) mbar_wait(&q_empty[t], (na - 1) & 1);
            mbar_arrive_expect_tx(&q_full[t], kAq3QBytes);
            tma_load_2d(sQ + t * kAq3QBytes, &map_q, &q_full[t], head * 64, row0 + t * 128);
            ++na;
          }
        }
        stamp(15, 0);
        for (int j = 1; j < nkv; ++j) load_kv(j);
        stamp(15, 0);
      }
      // every CTA has made its last fetch once all of them have been here: the last one rearms the counter for the next launch
      if (counters && atomicAdd(&counters[1], 1u) == gridDim.x - 1) {
        counters[0] = 0;
        counters[1] = 0;
        __threadfence();
      }
    }
  } else if (warp < 12) {
    // ===================================================== softmax groups (thread = query row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int t = warp >> 2, wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + t * kAq3TileCols;
    const uint32_t o_addr = s_addr + kAq3Keys;
    const float sl = p.scale_log2;
    int na = 0, ct = 0;
    float m_ref = -INFINITY, l_run = 0.f;

    auto tile = [&](auto nch_tag, auto full_tag, int j) {
      constexpr bool kFull = decltype(full_tag)::value;
      constexpr int nch = decltype(nch_tag)::value;
      const int nvalid = kFull ? kAq3Keys : p.ntok - j * kAq3Keys;
      uint32_t raw[3][32];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        if (ch < nch) tmem_ld_32x32b_x32(s_addr + ch * 32, raw[ch]);
      tmem_ld_wait();
      if (lane == 0) stamp(warp, 0);
      // ---- row maximum, four independent chains
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (ch < nch && (kFull || ch < nch - 1 || ch * 32 + i < nvalid)) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(raw[ch][i]));
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      // ---- lazy rescale: only when the maximum grew by more than 2^8 (always on the first tile of an item)
      const bool grow = (mx - m_ref) * sl > kAtcRescaleThreshold;
      // ---- P = exp2(S * sl - m * sl), packed to 16 bits in registers (the score registers die as we go)
      const float msl = (grow ? mx : m_ref) * sl;
      f32x2 rs2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[3][16];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (ch < nch) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const f32x2 xs = f2_fma(f2_pack(__uint_as_float(raw[ch][i]), __uint_as_float(raw[ch][i + 1])), f2_splat(sl), f2_splat(-msl));
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
            pk[ch][i >> 1] = Tr::pack2(p0, p1);
          }
        }
      }
      if (lane == 0) stamp(warp, 0);
      // ---- s_full(j) completed after P V(j-1) (issued before S(j)), so O may be rescaled without another wait; done here, with
      // the score registers dead, to stay inside the register budget
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
      // ---- P over the S it came from
      tmem_st_32x32b_x16(s_addr, pk[0]);
      if (nch > 1) tmem_st_32x32b_x16(s_addr + 16, pk[1]);
      if (nch > 2) tmem_st_32x32b_x16(s_addr + 32, pk[2]);
      tmem_st_wait();
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(rs2[0], rs2[1]), a0, a1);
        f2_unpack(f2_add(rs2[2], rs2[3]), b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
      tc_fence_before();            // TMEM stores (P, rescaled O) are ordered before the MMA that reads them
      mbar_arrive(&p_ready[t]);
      if (lane == 0) stamp(warp, 0);
    };

    using cuda::std::integral_constant;
    const int n_full = p.ntok / kAq3Keys;        // key tiles without a ragged end
    for (int n = 0;; ++n) {
      const int slot = n & 3;
      mbar_wait(&it_full[slot], (n >> 2) & 1);
      const int item = item_ring[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[slot]);
      if (item < 0) break;
      int img, head, g;
      decode(item, img, head, g);
      if (t >= q_tiles - 3 * g) continue;        // this slot has no query tile in the row's last group
      m_ref = -INFINITY;
      l_run = 0.f;
      for (int j = 0; j < n_full; ++j, ++ct) {
        mbar_wait(&s_full[t], ct & 1);
        tc_fence_after();
        if (lane == 0) stamp(warp, 0);
        tile(integral_constant<int, 3>{}, cuda::std::true_type{}, j);
      }
      if (n_full < nkv) {
        mbar_wait(&s_full[t], ct & 1);
        tc_fence_after();
        if (lane == 0) stamp(warp, 0);
        if (last_chunks == 1) tile(integral_constant<int, 1>{}, cuda::std::false_type{}, n_full);
        else if (last_chunks == 2) tile(integral_constant<int, 2>{}, cuda::std::false_type{}, n_full);
        else tile(integral_constant<int, 3>{}, cuda::std::false_type{}, n_full);
        ++ct;
      }
      // ---- normalise and store this row (128 contiguous bytes); O_t is handed back as soon as it is in registers
      mbar_wait(&o_full[t], na & 1);
      ++na;
      tc_fence_after();
      if (lane == 0) stamp(warp, 0);
      uint32_t o[2][32];
      tmem_ld_32x32b_x32(o_addr, o[0]);
      tmem_ld_32x32b_x32(o_addr + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_free[t]);
      if (lane == 0) stamp(warp, 0);
      // Thread = row: a warp-wide 16-byte global store touches 32 rows 2 KB apart, and even regrouped into whole rows the
      // eight stores of a warp took ~2000 clk in which the softmax warp issued nothing and the next item's first S waited
      // (clock64 trace, profiles/r02_attention_q3_traces.txt).  Instead: the 32 rows go into a 4 KB stage in the 128-byte
      // swizzle of the output's tensor map ({D, queries of the image, image}, 64 x 32 boxes: rows past the image's last
      // query are clipped by the copy engine) and ONE bulk tensor store takes them from there, asynchronously.
      const float inv = 1.0f / l_run;
      uint8_t* stage = sO + warp * (32 * 128);
      if (na > 1) {                           // the previous item's store has finished READING the stage
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = Tr::pack2(__uint_as_float(o[h][c * 8 + 0]) * inv, __uint_as_float(o[h][c * 8 + 1]) * inv);
          v.y = Tr::pack2(__uint_as_float(o[h][c * 8 + 2]) * inv, __uint_as_float(o[h][c * 8 + 3]) * inv);
          v.z = Tr::pack2(__uint_as_float(o[h][c * 8 + 4]) * inv, __uint_as_float(o[h][c * 8 + 5]) * inv);
          v.w = Tr::pack2(__uint_as_float(o[h][c * 8 + 6]) * inv, __uint_as_float(o[h][c * 8 + 7]) * inv);
          const int chunk = h * 4 + c;                                   // 16-byte chunk of this thread's row
          *reinterpret_cast<uint4*>(stage + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = v;
        }
      fence_proxy_async_smem();               // the generic-proxy writes above are visible to the bulk copy
      __syncwarp();
      if (lane == 0) stamp(warp, 0);
      if (lane == 0) {
        tma_store_3d(&map_out, stage, head * 64, (3 * g + t) * 128 + wq * 32, img);
        bulk_commit();
      }
      if (lane == 0) stamp(warp, 0);
    }
  } else {
    // ===================================================== MMA issuers: warp 13 + t owns query tile t
    // One thread per query tile, every wait blocking.  A single issuer for the three tiles was itself the bottleneck: polling
    // their barriers costs ~150 clk per test, and even in a fixed rotation the ~60 clk of instruction stream around each
    // tcgen05.mma (descriptor moves to uniform registers, the elect loop) add up to more than the tensor time of the 30 MMAs
    // of one round (profiles/r02_attention_q3_traces.txt).  Per key tile j:
    //     wait P_t(j);  O_t += P_t(j) V(j);  S_t(j + 1) = Q_t K(j + 1)^T
    // and after the last key tile straight on to S_t(0) of the next item, while the softmax warps normalise and store O_t.
    // A K / V ring stage goes back to the producer when all three issuers are through with it (barrier count 3); the issuer of
    // a query tile that the item does not have arrives by hand, paced by the stage's `full` barrier.
    // The whole warp runs the loop (every lane waits on the barriers) and one elected lane issues: with descriptors and
    // addresses computed from warp-uniform values the compiler keeps them in uniform registers, where tcgen05.mma wants them;
    // under `if (lane == 0)` it wraps every MMA in an elect loop with register -> uniform-register moves, ~110 clk per MMA.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    auto issuer = [&](auto t_tag) {
      constexpr int t = decltype(t_tag)::value;
      constexpr uint32_t idesc_o = umma_idesc_f16(Tr::kFmt, 128, 64) | (1u << 16);   // B (= V) is MN-major
      const uint32_t idesc_s_full = umma_idesc_f16(Tr::kFmt, 128, kAq3Keys);
      const uint32_t idesc_s_last = umma_idesc_f16(Tr::kFmt, 128, last_chunks * 32);
      const uint32_t tm = tmem_base + t * kAq3TileCols;
      const uint32_t smem0 = smem_u32(aq3_smem);
      const uint64_t qa = umma_desc_k_sw128(smem0 + t * kAq3QBytes);
      int na = 0;                          // items this query tile has taken part in (parity of q_full / o_full / o_free)
      int ct = 0;                          // key tiles it has been through (parity of p_ready)
      // the MMAs of S_t(cc) and what hangs on them; called by the elected lane
      auto issue_s = [&](int cc, bool last) {
        const int st = cc % kAq3Stages;
        const uint64_t b = umma_desc_k_sw128(smem0 + 3 * kAq3QBytes + st * kAq3KvBytes);
        const uint32_t idesc = last ? idesc_s_last : idesc_s_full;
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tm, qa + 2 * k, b + 2 * k, idesc, k != 0);
        tc_commit(&s_full[t]);
        stamp(12 + t, 0);
        if (last) tc_commit(&q_empty[t]);
        tc_commit(&k_empty[st]);
      };
      for (int n = 0;; ++n) {
        const int slot = n & 3;
        mbar_wait(&it_full[slot], (n >> 2) & 1);
        const int item = item_ring[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&it_empty[slot]);
        if (item < 0) break;
        int img_i, head_i, g_i;
        decode(item, img_i, head_i, g_i);
        const int nq = min(3, q_tiles - 3 * g_i);
        const int c0 = n * nkv;
        if (t >= nq) {
          for (int j = 0; j < nkv; ++j) {
            const int st = (c0 + j) % kAq3Stages;
            const uint32_t ph = ((c0 + j) / kAq3Stages) & 1;
            mbar_wait(&k_full[st], ph);
            if (lane == 0) mbar_arrive(&k_empty[st]);
            mbar_wait(&v_full[st], ph);
            if (lane == 0) mbar_arrive(&v_empty[st]);
          }
          continue;
        }
        mbar_wait(&k_full[c0 % kAq3Stages], (c0 / kAq3Stages) & 1);
        mbar_wait(&q_full[t], na & 1);
        tc_fence_after();
        if (elect_one()) issue_s(c0, nkv == 1);
        __syncwarp();
        for (int j = 0; j < nkv; ++j) {
          const int cc = c0 + j, st = cc % kAq3Stages;
          const bool last = j == nkv - 1;
          // K of the next S and V of this product arrived long ago: checked before the wait for P, so that the product and the
          // next S go out in one piece the moment P is there
          if (!last) mbar_wait(&k_full[(cc + 1) % kAq3Stages], ((cc + 1) / kAq3Stages) & 1);
          mbar_wait(&v_full[st], (cc / kAq3Stages) & 1);
          if (j == 0 && na > 0) mbar_wait(&o_free[t], (na - 1) & 1);
          mbar_wait(&p_ready[t], ct & 1);
          ++ct;
          tc_fence_after();
          const uint64_t vb = umma_desc_mn_sw128(smem0 + 3 * kAq3QBytes + (kAq3Stages + st) * kAq3KvBytes);
          const int ksteps = last ? 2 * last_chunks : 6;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 6; ++k)   // 16 keys per step: 8 packed P columns, two 8-row groups of V (2048 bytes)
              if (k < ksteps) tc_mma_f16_ts(tm + kAq3Keys, tm + 8 * k, vb + 128 * k, idesc_o, (j | k) != 0);
            stamp(12 + t, 0);
            if (last) tc_commit(&o_full[t]);
            tc_commit(&v_empty[st]);
            if (!last) issue_s(cc + 1, j + 1 == nkv - 1);
          }
          __syncwarp();
        }
        ++na;
      }
    };
    using cuda::std::integral_constant;
    if (warp == 13) issuer(integral_constant<int, 0>{});
    else if (warp == 14) issuer(integral_constant<int, 1>{});
    else issuer(integral_constant<int, 2>{});
  }

  if (warp < 12 && lane == 0) bulk_wait0();    // this warp's output stores have landed before its shared memory goes away
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mde
