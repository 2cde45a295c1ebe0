// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T      A, B 16-bit (fp16 or bf16) K-major, fp32 accumulation in TMEM.
//
// One CTA per SM loops over 128 x BLOCK_N output tiles.  Roles: warp 0 = TMA producer, warp 1 = MMA
// issuer (one thread), warp 2 = TMEM allocator, warps 4..11 = epilogue (TMEM -> registers -> fused
// epilogue -> global; two warps per TMEM lane quarter, each taking half of the tile's columns).  Two
// accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1; setmaxnreg moves
// the registers of the three single-thread roles to the epilogue warps.  Operand tiles are 128-byte-swizzled K-major boxes of 64 elements written by TMA.
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
  int tokens;           // ROW_TOKENS: T patch tokens per image
  int tok_skip;         // ROW_TOKENS: rows in front of the patch tokens in every image of the output (1 = cls; 1 + registers)
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
  int tma_out;          // plain row-major 16-bit output (bias / activation only): registers -> smem -> TMA store
  int gather_n;         // tma_out only: column boxes at or beyond gather_col0 are stored to gather_n tensors (the ranks'
  int gather_col0;      //   gathered K|V buffers, peer memory) instead of `out`; see GatherMaps
  int splits;           // split-K (tma_x only, >= 1): work item w = split * tiles + tile covers k-blocks [split * kb_per_split, ...);
  int kb_per_split;     //   the partial products meet in the L2's fp32 adds.  Fills the SMs when a small batch has few tiles
  int tma_x;            // x += gamma * (acc + bias) on plain rows as a bulk tensor reduction (fp32 add in the L2)
  // ---- fused depth head (BLOCK_N == 32 == N): z = sum_n relu(v_n) * head_w[n] + head_b
  const float* head_w;
  float head_b;
  float head_scale;     // metric: sigmoid(z) * head_scale; relative (head_scale < 0): relu(z)
  int head_act;         // 1: exp(z) (VGGT's depth head) instead of the two above
  float* head_out;      // [M] fp32
};

// kCtas == 2: a pair of CTAs on the two SMs of a TPC works on a 256 x BLOCK_N tile with tcgen05 cta_group::2: each CTA
// stages its own 128 rows of A but only HALF of the B tile, so the L2 -> shared-memory traffic per FLOP drops by a
// third (the single-CTA kernel sits at the ~64 B/clk/SM an SM can pull from L2) and two more stages fit.
template <int BLOCK_N, int kCtas = 1>
struct GemmCfg {
  static constexpr int kBlockM = 128;
  static constexpr int kBlockK = 64;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N / kCtas * kBlockK * 2;
  // k-blocks (of 64) per pipeline stage.  One thread issues every MMA: a barrier wait + four tcgen05.mma + a commit cost
  // about as many cycles as four 128 x 128 x 16 MMAs execute (measured: the tensor pipe of a 128-wide tile was 41 % busy
  // whatever the L2 traffic), so narrow tiles take two k-blocks per hand-shake.
  static constexpr int kKSub = BLOCK_N <= 128 ? 2 : 1;
  static constexpr int kStageBytes = (kABytes + kBBytes) * kKSub;
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiBytes = kEpiWarps * (32 * 32 * 4 + 32 * 4);   // per warp: 4 KB staging chunk + 32 row indices (a multiple of 1024)
  static constexpr int kMaxStages = (227 * 1024 - 1024 - 256 - kEpiBytes) / kStageBytes;   // 227 KB per CTA
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
  static constexpr int kThreads = 384;
};

// GELU(v) = v * Phi(v) with the exact-erf Phi, evaluated through Abramowitz & Stegun 7.1.26
// (|erf error| <= 1.5e-7): q = poly(t) * exp(-v^2/2), t = 1 / (1 + p |v| / sqrt(2)); Phi = v >= 0 ? 1 - q/2 : q/2.
// The negative branch has no cancellation; two MUFU ops (rcp, ex2) and about twelve FMA-class ops per element.
__device__ __forceinline__ float gelu_erf(float v) {
  const float a = fabsf(v);
  float t;   // rcp.approx (1 ulp) -- __frcp_rn expands to a ~60-instruction IEEE sequence
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(a, 0.3275911f * 0.70710678118654752440f, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * a * (-0.5f * 1.44269504088896340736f)));
  const float hq = 0.5f * poly * e;
  return v * (v >= 0.f ? 1.0f - hq : hq);
}

// Two GELUs at once on packed fp32 pairs, arranged so that sign handling costs nothing:
//   GELU(v) = v * Phi(v) = 0.5 * (v + |v| * erf(|v| / sqrt 2)),   erf(x) = 1 - poly(t) * exp(-x^2),  t = 1 / (1 + p x)
// i.e. den (FFMA2) - 2 rcp - poly - 2 ex2 - one FFMA2 for erf - one FFMA2 + one FMUL2 for the result; |v| rides on
// operand modifiers.  kShort picks Abramowitz & Stegun 7.1.25 (three terms, |erf error| <= 2.5e-5: relative error of
// the result <= 2.5e-5 for v > 0, absolute error <= 1.3e-5 |v| everywhere -- a twentieth of the fp16 rounding step
// 2.4e-4 |v| of the value that is stored, less still for bf16) instead of 7.1.26 (five terms, 1.5e-7).  About 13 / 15 issue
// slots per pair against 20 for the select-based form.  The 16-bit bulk-store epilogues use the short form for both
// operand types; the fp32 residual path (`gelu_erf`) keeps the five-term form.
template <bool kShort>
__device__ __forceinline__ void gelu_erf2(float& v0, float& v1) {
  const f32x2 v = f2_pack(v0, v1);
  const f32x2 a = f2_pack(fabsf(v0), fabsf(v1));
  constexpr float kP = kShort ? 0.47047f : 0.3275911f;
  const f32x2 den = f2_fma(a, f2_splat(kP * 0.70710678118654752440f), f2_splat(1.0f));
  float d0, d1, t0, t1;
  f2_unpack(den, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const f32x2 t = f2_pack(t0, t1);
  f32x2 poly;
  if (kShort) {
    poly = f2_fma(t, f2_splat(0.7478556f), f2_splat(-0.0958798f));
    poly = f2_fma(poly, t, f2_splat(0.3480242f));
  } else {
    poly = f2_fma(t, f2_splat(1.061405429f), f2_splat(-1.453152027f));
    poly = f2_fma(poly, t, f2_splat(1.421413741f));
    poly = f2_fma(poly, t, f2_splat(-0.284496736f));
    poly = f2_fma(poly, t, f2_splat(0.254829592f));
  }
  poly = f2_mul(poly, t);
  const f32x2 arg = f2_mul(f2_mul(a, f2_splat(-0.5f * 1.44269504088896340736f)), a);
  float g0, g1, e0, e1;
  f2_unpack(arg, g0, g1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const f32x2 erf_a = f2_fma(poly, f2_pack(-e0, -e1), f2_splat(1.0f));          // erf(|v| / sqrt 2) in [0, 1]
  f2_unpack(f2_mul(f2_fma(a, erf_a, v), f2_splat(0.5f)), v0, v1);
}

// GEMM -> all-gather in one kernel (sequence-sharded global attention: the QKV projection of this rank's tokens):
// m[r] describes THIS rank's row range inside rank r's gathered buffer ({columns, rows of this rank}, 64 x 32 boxes),
// so a box that hangs over this rank's last row is clipped instead of spilling into the next rank's rows.
struct GatherMaps {
  CUtensorMap m[8];
};

template <int BLOCK_N, typename T, int kCtas>
__global__ void __launch_bounds__(384, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_out, const __grid_constant__ GatherMaps gather, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, kCtas>;
  using Tr = F16Traits<T>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  constexpr int kKSub = Cfg::kKSub;
  uint8_t* smem_b = smem + kStages * kKSub * Cfg::kABytes;
  uint8_t* epi_smem = smem + kStages * Cfg::kStageBytes;   // per epilogue warp: 4 KB staging chunk (1024-byte aligned); then 8 x 32 row indices
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // A work item is one (128 * kCtas) x BLOCK_N tile; CTA `rank` of a pair owns rows [rank * 128, rank * 128 + 128) of it.
  const int rank = kCtas == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int num_tiles = ((p.m_tiles + kCtas - 1) / kCtas) * p.n_tiles;
  const int tile0 = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;
  auto m_block = [&](int tile) { return (tile / p.n_tiles) * kCtas + rank; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (p.tma_out || p.tma_x) prefetch_tmap(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], Cfg::kEpiWarps * kCtas);   // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kCtas == 2) { tmem_alloc_pair(tmem_slot, Cfg::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  griddep_launch_dependents();
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all();      // the peer's barriers are initialised before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                          // the prologue above overlapped the previous kernel's tail; its results are needed from here on

  if (warp == 0) {
    // ===================================================== TMA producer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      const uint32_t a_bytes = p.conv ? static_cast<uint32_t>(p.tile_w * p.tile_h * 128) : Cfg::kABytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int work = tile0; work < num_tiles * p.splits; work += tile_step) {
        const int tile = work % num_tiles;
        const int kb_begin = (work / num_tiles) * p.kb_per_split, kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
        const int m_blk = m_block(tile);
        const int n_blk = tile % p.n_tiles;
        int img = 0, y0 = 0, x0 = 0;
        if (p.conv) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_blk / per_img;
          const int t = m_blk % per_img;
          y0 = (t / p.tiles_x) * p.tile_h;
          x0 = (t % p.tiles_x) * p.tile_w;
        }
        for (int kb0 = kb_begin; kb0 < kb_end; kb0 += kKSub) {
          const int nsub = min(kKSub, kb_end - kb0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // pairs: both CTAs' bytes are counted on the leader's barrier; the leader alone arrives on it
          if (kCtas == 1 || rank == 0) mbar_arrive_expect_tx(&full_bar[stage], kCtas * nsub * (a_bytes + Cfg::kBBytes));
          for (int sub = 0; sub < nsub; ++sub) {
            const int kb = kb0 + sub;
            uint8_t* sa = smem_a + (stage * kKSub + sub) * Cfg::kABytes;
            uint8_t* sb = smem_b + (stage * kKSub + sub) * Cfg::kBBytes;
            if (kCtas == 2) {
              if (p.conv) {
                const int tap = kb / p.cin_blocks;
                const int c0 = (kb % p.cin_blocks) * 64;
                tma_load_4d_pair(sa, &map_a, &full_bar[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img);
              } else {
                tma_load_2d_pair(sa, &map_a, &full_bar[stage], kb * 64, m_blk * 128);
              }
              tma_load_2d_pair(sb, &map_b, &full_bar[stage], kb * 64, n_blk * BLOCK_N + rank * (BLOCK_N / 2));
            } else {
              if (p.conv) {
                const int tap = kb / p.cin_blocks;
                const int c0 = (kb % p.cin_blocks) * 64;
                tma_load_4d(sa, &map_a, &full_bar[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img);
              } else {
                tma_load_2d(sa, &map_a, &full_bar[stage], kb * 64, m_blk * 128);
              }
              tma_load_2d(sb, &map_b, &full_bar[stage], kb * 64, n_blk * BLOCK_N);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // The whole warp runs the loop and one elected lane issues: with the descriptors computed from warp-uniform values the
    // compiler keeps them in uniform registers and issues the four MMAs of a k block back to back.  Under `if (lane == 0)` it
    // wrapped every tcgen05.mma in an elect loop with register -> uniform-register moves, ~110 clk per MMA -- more than the
    // 64 clk a 128 x 128 x 16 MMA takes, so the single-CTA N = 128 tiles were issue-bound (profiles/r02_attention_q3_traces.txt).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(Tr::kFmt, 128 * kCtas, BLOCK_N);
      const uint32_t smem_a0 = smem_u32(smem_a), smem_b0 = smem_u32(smem_b);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = tile0; work < num_tiles * p.splits; work += tile_step) {
        const int kb_begin = (work / num_tiles) * p.kb_per_split, kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb0 = kb_begin; kb0 < kb_end; kb0 += kKSub) {
          const int nsub = min(kKSub, kb_end - kb0);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            for (int sub = 0; sub < nsub; ++sub) {
              const uint64_t a_desc = umma_desc_k_sw128(smem_a0 + (stage * kKSub + sub) * Cfg::kABytes);
              const uint64_t b_desc = umma_desc_k_sw128(smem_b0 + (stage * kKSub + sub) * Cfg::kBBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // +32 bytes per UMMA_K=16 step inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
                const uint32_t accumulate = ((kb0 + sub - kb_begin) | k) != 0;
                if (kCtas == 2) tc_mma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accumulate);
                else tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accumulate);
              }
            }
            if (kCtas == 2) tc_commit_pair(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (kCtas == 2) tc_commit_pair(&tmem_full[acc]); else tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // TMEM allocator warp and the spare one
  } else {
    // ===================================================== epilogue
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // Phase 1 (thread = row, as tcgen05.ld 32x32b delivers it): a raw 32 x 32 fp32 chunk goes to this
    // warp's swizzled staging buffer; the TMEM load of the next chunk is already in flight.
    // Phase 2 (lanes = columns): bias / activation / LayerScale / residuals and all global traffic run
    // over whole row segments, so every 32-byte sector that is touched is touched completely, per-column
    // parameters live in one register per lane, and all loads of a chunk are issued before its stores.
    const int ew = warp - 4;                 // 0..7
    const int quarter = ew & 3;              // == warp % 4: the TMEM lane quarter this warp may read
    const int half = ew >> 2;                // which half of the tile's column chunks this warp takes
    const int r = quarter * 32 + lane;       // row inside the 128-row tile
    float* stage_f = reinterpret_cast<float*>(epi_smem) + ew * (32 * 32);
    int* row_off = reinterpret_cast<int*>(epi_smem + Cfg::kEpiWarps * 32 * 32 * 4) + ew * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    // hand an accumulator stage back to the MMA issuer (which lives in the pair's leader CTA)
    auto release_acc = [&](int a) {
      if (kCtas == 2 && rank != 0) mbar_arrive_cluster(&tmem_empty[a], 0);
      else mbar_arrive(&tmem_empty[a]);
    };
    T* out = static_cast<T*>(p.out);
    T* out_relu = static_cast<T*>(p.out_relu);
    const T* res1 = static_cast<const T*>(p.res1);
    const T* res2 = static_cast<const T*>(p.res2);
    if (p.tma_out) {
      // ---- plain row-major 16-bit output (QKV, FC1, DPT projections): thread = row, all arithmetic on the
      // registers tcgen05.ld delivered, two 32-column chunks packed into one 128-byte-swizzled 32 x 64 box
      // and written by one bulk tensor store per warp (the tensor map clips rows >= M).  About a third of
      // the instructions of the staged path below: nothing is re-read from shared memory, no addresses,
      // no predicates.
      uint8_t* buf = epi_smem + ew * 4096;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_blk = m_block(tile);
        const int n_tile = (tile % p.n_tiles) * BLOCK_N;
        constexpr int kPerWarp = BLOCK_N / 64;             // chunks per warp; the host guarantees N % 32 == 0, BLOCK_N >= 128 (a last tile
                                                           // that N does not fill: B rows read as zeros, store boxes clipped)
        const int ch_begin = half * kPerWarp;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
        uint32_t raw[32];
        tmem_ld_32x32b_x32(t_base + ch_begin * 32, raw);
#pragma unroll 1
        for (int c = 0; c < kPerWarp; ++c) {
          const int n_base = n_tile + (ch_begin + c) * 32;
          float4 bia[8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            bia[g] = (p.bias && n_base < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + n_base) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (c + 1 < kPerWarp) {
            tmem_ld_32x32b_x32(t_base + (ch_begin + c + 1) * 32, raw);   // in flight during the arithmetic
          } else {
            tc_fence_before();                 // every TMEM read of this stage is complete: hand it back now
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            f32x2 lo = f2_add(f2_pack(v[4 * g], v[4 * g + 1]), f2_pack(bia[g].x, bia[g].y));
            f32x2 hi = f2_add(f2_pack(v[4 * g + 2], v[4 * g + 3]), f2_pack(bia[g].z, bia[g].w));
            f2_unpack(lo, v[4 * g], v[4 * g + 1]);
            f2_unpack(hi, v[4 * g + 2], v[4 * g + 3]);
          }
          if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_erf2<true>(v[j], v[j + 1]);
          } else if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if ((c & 1) == 0) {
            if (lane == 0) bulk_wait_read0();  // the previous store has finished reading the staging box
            __syncwarp();
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = Tr::pack2(v[8 * g], v[8 * g + 1]); u.y = Tr::pack2(v[8 * g + 2], v[8 * g + 3]);
            u.z = Tr::pack2(v[8 * g + 4], v[8 * g + 5]); u.w = Tr::pack2(v[8 * g + 6], v[8 * g + 7]);
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((((c & 1) * 4 + g) ^ (lane & 7)) << 4)) = u;
          }
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.gather_n > 0 && n_base - 32 >= p.gather_col0) {
                // K / V columns: one bulk store per rank, into local HBM for our own buffer and over NVLink for the peers'
                for (int r = 0; r < p.gather_n; ++r)
                  tma_store_2d(&gather.m[r], buf, n_base - 32 - p.gather_col0, m_blk * 128 + quarter * 32);
              } else {
                tma_store_2d(&map_out, buf, n_base - 32, m_blk * 128 + quarter * 32);
              }
              bulk_commit();
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) bulk_wait0();
    } else if (p.tma_x) {
      // ---- residual-stream update on plain rows (attention projection, FC2) as a bulk tensor REDUCTION:
      // x[box] += gamma * (acc + bias), the add done by the L2 (cp.reduce.async.bulk.tensor ... .add, fp32).  The SM never
      // loads x: nothing to wait for, no global load/store instructions, addresses or predicates in the loop; rows >= M are
      // clipped.  (A TMA-load / update-in-place / TMA-store variant measured 897 vs 990 TFLOP/s on the projection.)
      uint8_t* buf = epi_smem + ew * 4096;
      constexpr int kPerWarp = BLOCK_N / 64;
      const int ch_begin = half * kPerWarp;
      for (int work = tile0; work < num_tiles * p.splits; work += tile_step) {
        const int tile = work % num_tiles;
        const bool first_split = work < num_tiles;       // the bias is added once per output element
        const int row0 = m_block(tile) * 128 + quarter * 32;
        const int n_tile = (tile % p.n_tiles) * BLOCK_N;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
        uint32_t raw[32];
        tmem_ld_32x32b_x32(t_base + ch_begin * 32, raw);
#pragma unroll 1
        for (int c = 0; c < kPerWarp; ++c) {
          const int n_base = n_tile + (ch_begin + c) * 32;
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (c + 1 < kPerWarp) {
            tmem_ld_32x32b_x32(t_base + (ch_begin + c + 1) * 32, raw);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
          }
          if (lane == 0) bulk_wait_read0();      // the previous reduction has drained the staging box
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 bia = make_float4(0.f, 0.f, 0.f, 0.f), gam = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.bias && first_split) bia = __ldg(reinterpret_cast<const float4*>(p.bias + n_base) + g);
            if (p.gamma) gam = __ldg(reinterpret_cast<const float4*>(p.gamma + n_base) + g);
            float4 d;
            d.x = (v[4 * g] + bia.x) * gam.x; d.y = (v[4 * g + 1] + bia.y) * gam.y;
            d.z = (v[4 * g + 2] + bia.z) * gam.z; d.w = (v[4 * g + 3] + bia.w) * gam.w;
            *reinterpret_cast<float4*>(buf + lane * 128 + ((g ^ (lane & 7)) << 4)) = d;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&map_out, buf, n_base, row0);
            bulk_commit();
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) bulk_wait0();
    } else
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = m_block(tile);
      const int n_blk = tile % p.n_tiles;
      // ---- where does this thread's row go?
      bool valid;
      long long orow;            // output row (ROW_SHUFFLE: row of sub-pixel (0,0))
      if (p.conv) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int img = m_blk / per_img;
        const int t = m_blk % per_img;
        const int y = (t / p.tiles_x) * p.tile_h + r / p.tile_w;
        const int x = (t % p.tiles_x) * p.tile_w + r % p.tile_w;
        valid = (r < p.tile_w * p.tile_h) && y < p.H && x < p.W && m_blk < p.m_tiles;
        orow = (static_cast<long long>(img) * p.H + y) * p.W + x;
      } else {
        const long long m = static_cast<long long>(m_blk) * 128 + r;
        valid = m < p.M;
        orow = m;
        if (p.row_map == ROW_TOKENS) {
          orow = (m / p.tokens) * (p.tokens + p.tok_skip) + p.tok_skip + (m % p.tokens);
        } else if (p.row_map == ROW_SHUFFLE) {
          const int hw = p.H * p.W;
          const long long b = m / hw;
          const int rem = static_cast<int>(m % hw);
          const int y = rem / p.W, x = rem % p.W;
          orow = (b * (p.H * p.shuffle_s) + static_cast<long long>(y) * p.shuffle_s) * (p.W * p.shuffle_s) +
                 static_cast<long long>(x) * p.shuffle_s;
        }
      }
      __syncwarp();                          // previous tile's phase 2 is done with row_off
      row_off[lane] = valid ? static_cast<int>(orow) : -1;   // output rows fit 31 bits (checked on the host)
      const int n_tile = n_blk * BLOCK_N;
      const int chunks = min(BLOCK_N, p.N - n_tile + 31) / 32;   // warp-uniform; N is a multiple of 8
      const int ch_begin = half == 0 ? 0 : (chunks + 1) / 2;
      const int ch_end = half == 0 ? (chunks + 1) / 2 : chunks;
      // The residual stream values this warp will read-modify-write after its first chunk: pull them into L2
      // while the MMAs of the tile are still running (one 128-byte line per row and chunk).
      if (p.x && p.accumulate_x && valid) {
        const char* xrow = reinterpret_cast<const char*>(p.x + orow * p.ld_out + n_tile);
        for (int i = ch_begin + 1; i < ch_end; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow + i * 128));
      }
      // fp32 pass (2a): this lane's 8 rows and the residual-stream values (or pos-embed rows) it will add.
      // Nobody else touches these elements, so chunk c+1 is requested as soon as chunk c has been consumed
      // and the first chunk before the accumulator is even complete: the loads overlap the MMAs / phase 1.
      __syncwarp();
      int rx[8];
      float4 xin[8];
      auto fetch_x = [&](int ch) {
        const int n = n_tile + ch * 32 + 4 * (lane & 7);
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
          xin[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rx[pass] >= 0 && n < p.N) {
            if (p.accumulate_x) xin[pass] = *reinterpret_cast<const float4*>(p.x + static_cast<long long>(rx[pass]) * p.ld_out + n);
            else if (p.pos) xin[pass] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<long long>(rx[pass] % (p.tokens + p.tok_skip) - p.tok_skip + 1) * p.ld_out + n));
          }
        }
      };
      if (p.x) {
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) rx[pass] = row_off[pass * 4 + (lane >> 3)];
        if (ch_begin < ch_end) fetch_x(ch_begin);
      }

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;

      uint32_t raw[32];
      if (ch_begin < ch_end) tmem_ld_32x32b_x32(t_base + ch_begin * 32, raw);
#pragma unroll 1
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int n_base = n_tile + ch * 32;
        // Per-lane column parameters of this chunk, requested before the TMEM wait so their latency hides
        // behind it: lanes own 4 columns in the fp32 pass (2a) and 8 columns in the 16-bit pass (2b).
        float4 bia_a = make_float4(0.f, 0.f, 0.f, 0.f), gam_a = make_float4(1.f, 1.f, 1.f, 1.f);
        float4 bia_b0 = bia_a, bia_b1 = bia_a, gam_b0 = gam_a, gam_b1 = gam_a;
        if (p.head_w == nullptr) {
          const int na = n_base + 4 * (lane & 7);
          if (p.x && na < p.N) {
            if (p.bias) bia_a = __ldg(reinterpret_cast<const float4*>(p.bias + na));
            if (p.gamma) gam_a = __ldg(reinterpret_cast<const float4*>(p.gamma + na));
          }
          const int nb = n_base + 8 * (lane & 3);
          if ((out || out_relu) && nb < p.N) {
            const int pc = p.row_map == ROW_SHUFFLE ? nb % p.shuffle_cout : nb;
            if (p.bias) {
              bia_b0 = __ldg(reinterpret_cast<const float4*>(p.bias + pc));
              bia_b1 = __ldg(reinterpret_cast<const float4*>(p.bias + pc + 4));
            }
            if (p.gamma) {
              gam_b0 = __ldg(reinterpret_cast<const float4*>(p.gamma + pc));
              gam_b1 = __ldg(reinterpret_cast<const float4*>(p.gamma + pc + 4));
            }
          }
        }
        tmem_ld_wait();
        if (p.head_w != nullptr) {
          // fused depth head: 3x3 conv (+bias, ReLU) -> 1x1 conv 32->1 -> activation
          float z = p.head_b;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(raw[j]) + __ldg(p.bias + n_base + j);
            z = fmaf(fmaxf(v, 0.f), __ldg(p.head_w + n_base + j), z);
          }
          if (valid) p.head_out[orow] = p.head_act == 1 ? expf(z) : (p.head_scale < 0.f ? fmaxf(z, 0.f) : p.head_scale / (1.0f + __expf(-z)));
          continue;
        }
        // ---- phase 1
        __syncwarp();                        // previous chunk's phase 2 has drained the staging buffer
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(stage_f + lane * 32 + ((g ^ (lane & 7)) << 2)) =
              make_uint4(raw[4 * g], raw[4 * g + 1], raw[4 * g + 2], raw[4 * g + 3]);
        if (ch + 1 < ch_end) tmem_ld_32x32b_x32(t_base + (ch + 1) * 32, raw);   // in flight during phase 2
        __syncwarp();
        // ---- phase 2a: fp32 residual stream, 8 lanes x float4 per row, 4 rows per pass
        if (p.x) {
          const int c4 = lane & 7;
          const int n = n_base + 4 * c4;
          if (n < p.N) {
            const float4 bia = bia_a, gam = gam_a;
#pragma unroll
            for (int pass = 0; pass < 8; ++pass) {
              if (rx[pass] < 0) continue;
              const int rr = pass * 4 + (lane >> 3);
              float4 v = *reinterpret_cast<const float4*>(stage_f + rr * 32 + ((c4 ^ (rr & 7)) << 2));
              v.x += bia.x; v.y += bia.y; v.z += bia.z; v.w += bia.w;
              if (p.act == ACT_GELU) {
                v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
              } else if (p.act == ACT_RELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              }
              v.x = fmaf(v.x, gam.x, xin[pass].x); v.y = fmaf(v.y, gam.y, xin[pass].y);
              v.z = fmaf(v.z, gam.z, xin[pass].z); v.w = fmaf(v.w, gam.w, xin[pass].w);
              *reinterpret_cast<float4*>(p.x + static_cast<long long>(rx[pass]) * p.ld_out + n) = v;
            }
          }
          if (ch + 1 < ch_end) fetch_x(ch + 1);
        }
        // ---- phase 2b: 16-bit outputs (+ 16-bit residuals), 4 lanes x 8 columns per row, 8 rows per pass
        if (out || out_relu) {
          const int c8 = lane & 3;
          const int n = n_base + 8 * c8;
          if (n < p.N) {
            int col = n;
            long long sub = 0;
            if (p.row_map == ROW_SHUFFLE) {
              const int q = n / p.shuffle_cout;
              col = n % p.shuffle_cout;
              sub = static_cast<long long>(q / p.shuffle_s) * (p.W * p.shuffle_s) + (q % p.shuffle_s);
            }
            const float bia[8] = {bia_b0.x, bia_b0.y, bia_b0.z, bia_b0.w, bia_b1.x, bia_b1.y, bia_b1.z, bia_b1.w};
            const float gam[8] = {gam_b0.x, gam_b0.y, gam_b0.z, gam_b0.w, gam_b1.x, gam_b1.y, gam_b1.z, gam_b1.w};
            long long off[4];
            uint4 ra[4], rb[4];
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const long long ro = static_cast<long long>(row_off[pass * 8 + (lane >> 2)]);
              off[pass] = ro < 0 ? -1 : (ro + sub) * p.ld_out + col;
              ra[pass] = make_uint4(0u, 0u, 0u, 0u);
              rb[pass] = make_uint4(0u, 0u, 0u, 0u);
              if (off[pass] >= 0) {
                if (res1) ra[pass] = *reinterpret_cast<const uint4*>(res1 + off[pass]);
                if (res2) rb[pass] = *reinterpret_cast<const uint4*>(res2 + off[pass]);
              }
            }
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              if (off[pass] < 0) continue;
              const int rr = pass * 8 + (lane >> 2);
              const float4 v0 = *reinterpret_cast<const float4*>(stage_f + rr * 32 + (((2 * c8) ^ (rr & 7)) << 2));
              const float4 v1 = *reinterpret_cast<const float4*>(stage_f + rr * 32 + (((2 * c8 + 1) ^ (rr & 7)) << 2));
              float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
              const uint32_t* pa = &ra[pass].x;
              const uint32_t* pb = &rb[pass].x;
              // every branch below is uniform over the launch: untaken ones cost nothing per element
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += bia[j];
              }
              if (p.act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) gelu_erf2<true>(v[j], v[j + 1]);
              } else if (p.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (p.gamma) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= gam[j];
              }
              if (res1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 a = Tr::unpack2(pa[j]);
                  v[2 * j] += a.x; v[2 * j + 1] += a.y;
                }
              }
              if (res2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 b = Tr::unpack2(pb[j]);
                  v[2 * j] += b.x; v[2 * j + 1] += b.y;
                }
              }
              if (out) {
                uint4 u;
                u.x = Tr::pack2(v[0], v[1]); u.y = Tr::pack2(v[2], v[3]);
                u.z = Tr::pack2(v[4], v[5]); u.w = Tr::pack2(v[6], v[7]);
                *reinterpret_cast<uint4*>(out + off[pass]) = u;
              }
              if (out_relu) {
                uint4 u;
                u.x = Tr::pack2(fmaxf(v[0], 0.f), fmaxf(v[1], 0.f)); u.y = Tr::pack2(fmaxf(v[2], 0.f), fmaxf(v[3], 0.f));
                u.z = Tr::pack2(fmaxf(v[4], 0.f), fmaxf(v[5], 0.f)); u.w = Tr::pack2(fmaxf(v[6], 0.f), fmaxf(v[7], 0.f));
                *reinterpret_cast<uint4*>(out_relu + off[pass]) = u;
              }
            }
          }
        }
      }
      // all of this warp's TMEM reads of the stage are complete (wait::ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all();      // nothing of the pair is still addressed to this CTA
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (kCtas == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mde
