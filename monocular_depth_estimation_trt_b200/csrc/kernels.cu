#include <utility>
// Kernel instantiations, launchers and the single-kernel C-ABI entry points (mde_k_*).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <set>
#include <vector>

#include "attention_mma.cuh"
#include "attention_tc.cuh"
#include "attention_q3.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "host_common.h"
#include "upconv_head.cuh"

namespace mde {

// ------------------------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
void clear_error() { g_last_error.clear(); }
const char* last_error_cstr() { return g_last_error.c_str(); }

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs).
int num_sms() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] <= 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

// Function attributes (dynamic shared memory limit, carve-out) are per DEVICE: set once per (kernel, device ordinal).
static int ensure_func_attrs(const void* kern, int smem_bytes, bool max_carveout) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  MDE_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({kern, dev})) return MDE_OK;
  MDE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (max_carveout) MDE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  done.insert({kern, dev});
  return MDE_OK;
}

// Launch options of the calling thread, set by the engine around an enqueue from its description (mde_engine_desc.flags);
// the single-kernel entry points run with the defaults.
static thread_local LaunchOpts g_opts;
void set_launch_opts(const LaunchOpts& o) { g_opts = o; }
LaunchOpts launch_opts() { return g_opts; }

// ------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MDE_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p)
      return fail(MDE_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return MDE_OK;
}

static int encode_map(CUtensorMap* map, int precision, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box);
static int get_encode(EncodeTiledFn* fn);

// The bulk-store epilogue serves plain row-major 16-bit outputs: bias and activation only, whole tiles in N.
static int maybe_tma_out(GemmOp* op, int precision, const mde_epilogue* ep, long long m, int n) {
  memset(&op->map_out, 0, sizeof(op->map_out));
  memset(&op->gather, 0, sizeof(op->gather));
  GemmParams& p = op->p;
  p.tma_out = 0;
  p.tma_x = 0;
  p.gather_n = 0;
  p.gather_col0 = 0;
  if (ep->gather_n < 0 || ep->gather_n > 8) return fail(MDE_ERR_INVALID, "gemm: at most 8 gather destinations");
  // residual-stream update on plain rows: fp32 boxes of gamma * (acc + bias) are added to x by bulk tensor reductions
  if (ep->d_x && ep->accumulate_x && !ep->d_out && !ep->d_out_relu && !ep->d_res1 && !ep->d_res2 && !ep->d_pos && !ep->d_head_w &&
      ep->act == 0 && !p.conv && p.row_map == ROW_IDENTITY && op->block_n >= 128 && n % op->block_n == 0 && ep->ld_out % 4 == 0 &&
      !(reinterpret_cast<uintptr_t>(ep->d_x) & 15)) {
    EncodeTiledFn enc;
    MDE_TRY(get_encode(&enc));
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(m)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(ep->ld_out) * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&op->map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ep->d_x, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MDE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the fp32 residual stream", static_cast<int>(r));
    p.tma_x = 1;
    return MDE_OK;
  }
  // N need not fill the last tile: B rows beyond N are zero-filled by the tensor map, store boxes beyond N are clipped (whole
  // 32-column chunks only, so that the per-chunk bias loads stay inside the vector)
  if (!ep->d_out || ep->d_x || ep->d_res1 || ep->d_res2 || ep->d_out_relu || ep->d_gamma || ep->d_head_w || p.conv ||
      p.row_map != ROW_IDENTITY || op->block_n < 128 || n % 32 || (n % op->block_n && ep->gather_n > 0) ||
      (reinterpret_cast<uintptr_t>(ep->d_out) & 15))
    return MDE_OK;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(m)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(ep->ld_out) * 2};
  cuuint32_t box[2] = {64, 32};
  MDE_TRY(encode_map(&op->map_out, precision, ep->d_out, 2, dims, str, box));
  p.tma_out = 1;
  if (ep->gather_n > 0) {
    if (ep->gather_col0 % 64 || ep->gather_col0 < 0 || ep->gather_col0 >= n || ep->gather_ld % 8 || ep->gather_ld < n - ep->gather_col0)
      return fail(MDE_ERR_INVALID, "gemm: gather_col0 must be a multiple of 64 inside [0, N) and gather_ld >= N - gather_col0, a multiple of 8");
    for (int r = 0; r < ep->gather_n; ++r) {
      if (!ep->d_gather[r] || (reinterpret_cast<uintptr_t>(ep->d_gather[r]) & 15)) return fail(MDE_ERR_INVALID, "gemm: gather destination %d is null or misaligned", r);
      cuuint64_t gdims[2] = {static_cast<cuuint64_t>(n - ep->gather_col0), static_cast<cuuint64_t>(m)};
      cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ep->gather_ld) * 2};
      MDE_TRY(encode_map(&op->gather.m[r], precision, ep->d_gather[r], 2, gdims, gstr, box));
    }
    p.gather_n = ep->gather_n;
    p.gather_col0 = ep->gather_col0;
  }
  return MDE_OK;
}

static int encode_map(CUtensorMap* map, int precision, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn enc;
  MDE_TRY(get_encode(&enc));
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = precision == MDE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MDE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,...] box [%u,%u,...]",
                static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return MDE_OK;
}

static int pick_block_n(int n) {
  // smallest padded N wins; ties go to the wider tile (fewer A re-reads, longer MMA bursts).  From 128 columns on only the
  // wide tiles compete: a narrow tile re-reads its 128-wide A operand per 32 / 64 columns and has no bulk-store epilogue, which
  // costs more than the zero columns of a partly filled wide tile (N = 288: 128-wide tiles, the third one a quarter full).
  const int cands[4] = {256, 128, 64, 32};
  int best = 32;
  long long best_pad = 1LL << 60;
  for (int bn : cands) {
    if (n >= 128 && bn < 128) continue;
    const long long pad = static_cast<long long>((n + bn - 1) / bn) * bn;
    if (pad < best_pad) { best_pad = pad; best = bn; }
  }
  return best;
}

// One persistent CTA per SM -- or, for wide tiles with enough rows, one CTA pair per TPC.
static void pick_ctas(GemmOp* op) {
  // pairs pay off for 256-wide tiles with a real K loop; short K (a few k-blocks per tile) is epilogue-bound either way
  const bool wide = op->block_n == 256 || (op->block_n == 128 && op->p.num_k_blocks >= 8);
  op->ctas = (wide && op->p.m_tiles >= 2 && op->p.num_k_blocks >= 4) ? 2 : 1;
}

// Tile width and CTA pairing.  A problem that fills the chip several times over takes the widest tile (fewest operand
// re-reads, longest MMA bursts: pick_block_n / pick_ctas).  A SMALL problem -- batch 1: 11 row tiles -- is latency-bound: with
// 256-wide pair tiles FC2 (N = 1024) runs on 48 of 148 SMs for 64 k-blocks.  There the candidates are scored by
//     waves x (k-blocks x 4 MMAs x BLOCK_N / 2 clk  +  epilogue  +  fixed cost per tile)
// and narrower tiles win when they put more SMs to work or cut the last partial wave (FC1: 96 pair tiles on 74 pairs).
// p.m_tiles and p.num_k_blocks are set; sets op->block_n, op->ctas, p.n_tiles.
// mde_k_gemm_tiled (a tuning probe, tools/gemm_tiling_probe.py): the next GEMM op of this thread takes the given tile width,
// pairing and split count instead of the picked ones
struct TilingOverride { int block_n = 0, ctas = 0, splits = 0; };
static thread_local TilingOverride g_tiling;

static void pick_tiling(GemmOp* op, int n, bool fused_head, bool whole_wide_tiles = false) {
  GemmParams& p = op->p;
  if (g_tiling.block_n > 0 && !fused_head) {
    op->block_n = g_tiling.block_n;
    op->ctas = g_tiling.ctas == 2 && g_tiling.block_n >= 128 && p.m_tiles >= 2 ? 2 : 1;
    p.n_tiles = (n + op->block_n - 1) / op->block_n;
    return;
  }
  op->block_n = fused_head ? 32 : pick_block_n(n);
  p.n_tiles = (n + op->block_n - 1) / op->block_n;
  pick_ctas(op);
  const int sms = num_sms();
  if (fused_head || n < 128 || sms <= 0 || g_opts.split_k) return;    // split-K engines keep the wide tiles and split those
  {
    const int units = op->ctas == 2 ? sms / 2 : sms;
    const long long tiles = op->ctas == 2 ? static_cast<long long>((p.m_tiles + 1) / 2) * p.n_tiles : static_cast<long long>(p.m_tiles) * p.n_tiles;
    if (tiles >= 2LL * units) return;
  }
  long long best_cost = 1LL << 60, best_work = 1LL << 60;
  int best_bn = op->block_n, best_ctas = op->ctas;
  const int bns[3] = {256, 128, 64};
  for (int bn : bns) {
    for (int ctas = 2; ctas >= 1; --ctas) {
      if (ctas == 2 && (bn < 128 || p.m_tiles < 2 || p.num_k_blocks < 4)) continue;
      if (whole_wide_tiles && (bn < 128 || n % bn)) continue;     // the fused gather stores whole 128- / 256-column tiles
      const int n_tiles = (n + bn - 1) / bn;
      const long long m_units = ctas == 2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
      const long long tiles = m_units * n_tiles;
      const int units = sms / ctas;
      const long long waves = (tiles + units - 1) / units;
      const long long per_tile = static_cast<long long>(p.num_k_blocks) * 2 * bn + (bn >= 128 ? 10 : 30) * bn + 1500;
      const long long cost = waves * per_tile;
      const long long work = m_units * ctas * n_tiles * bn;       // padded rows x columns (in 128-row units)
      if (cost < best_cost || (cost == best_cost && work < best_work)) {
        best_cost = cost; best_work = work; best_bn = bn; best_ctas = ctas;
      }
    }
  }
  op->block_n = best_bn;
  op->ctas = best_ctas;
  p.n_tiles = (n + best_bn - 1) / best_bn;
}
static int pick_grid(GemmOp* op) {
  const int sms = num_sms();
  if (sms <= 0) return fail(MDE_ERR_CUDA, "no CUDA device");
  const GemmParams& p = op->p;
  GemmParams& pw = op->p;
  pw.splits = 1;
  pw.kb_per_split = p.num_k_blocks;
  const int units = op->ctas == 2 ? sms / 2 : sms;                                        // CTAs or CTA pairs that can run at once
  const int tiles = op->ctas == 2 ? ((p.m_tiles + 1) / 2) * p.n_tiles : p.m_tiles * p.n_tiles;
  if (g_tiling.splits > 1 && p.tma_x) {
    pw.kb_per_split = (p.num_k_blocks + g_tiling.splits - 1) / g_tiling.splits;
    pw.splits = (p.num_k_blocks + pw.kb_per_split - 1) / pw.kb_per_split;
  } else if (p.tma_x && tiles * 2 <= units && p.num_k_blocks >= 16 && g_opts.split_k) {
    // small batch: a handful of tiles on 148 SMs.  Split K so that every SM gets a piece (at least 8 k-blocks each);
    // the reduction epilogue adds the partial products in the L2.  Opt-in (MDE_FLAG_SPLIT_K in mde_engine_desc.flags, part
    // of the engine fingerprint): the fp32 adds happen in arrival order, so results are no longer reproducible bit for bit.
    int splits = std::min(std::min(units / tiles, p.num_k_blocks / 8), 8);
    pw.kb_per_split = (p.num_k_blocks + splits - 1) / splits;
    pw.splits = (p.num_k_blocks + pw.kb_per_split - 1) / pw.kb_per_split;
  }
  op->grid = op->ctas * std::min(units, tiles * pw.splits);
  return MDE_OK;
}

static int fill_epilogue(GemmParams& p, const mde_epilogue* ep, long long rows_out) {
  p.row_map = ROW_IDENTITY;
  p.tokens = 0; p.tok_skip = 1; p.shuffle_s = 0; p.shuffle_cout = 0;
  p.act = ep->act;
  p.ld_out = ep->ld_out;
  p.accumulate_x = ep->accumulate_x;
  p.bias = ep->d_bias; p.gamma = ep->d_gamma; p.pos = ep->d_pos;
  p.x = ep->d_x; p.res1 = ep->d_res1; p.res2 = ep->d_res2; p.out = ep->d_out; p.out_relu = ep->d_out_relu;
  p.head_w = ep->d_head_w; p.head_b = ep->head_b;
  p.head_scale = ep->head_scale > 0.f ? ep->head_scale : -1.f;
  p.head_act = ep->head_act;
  if (ep->head_act != 0 && ep->head_act != 1) return fail(MDE_ERR_INVALID, "unknown head activation %d", ep->head_act);
  if (ep->token_skip < 0) return fail(MDE_ERR_INVALID, "token_skip must not be negative");
  p.head_out = ep->d_head_out;
  if (ep->ld_out % 8 != 0 && !ep->d_head_w) return fail(MDE_ERR_INVALID, "ld_out must be a multiple of 8");
  if (rows_out > 0x7fffffffLL) return fail(MDE_ERR_INVALID, "more than 2^31 output rows");
  if (ep->act < 0 || ep->act > 2) return fail(MDE_ERR_INVALID, "unknown activation %d", ep->act);
  return MDE_OK;
}

int make_gemm_op(GemmOp* op, int precision, const void* d_a, long long m, int k, int lda, const void* d_b, int n,
                 int ldb, const mde_epilogue* ep) {
  if (precision != MDE_FP16 && precision != MDE_BF16) return fail(MDE_ERR_INVALID, "precision must be MDE_FP16 or MDE_BF16");
  if (m <= 0 || n <= 0 || k <= 0) return fail(MDE_ERR_INVALID, "gemm: empty problem m=%lld n=%d k=%d", m, n, k);
  if (lda % 8 || ldb % 8 || n % 8) return fail(MDE_ERR_INVALID, "gemm: lda, ldb and n must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(d_a) | reinterpret_cast<uintptr_t>(d_b)) & 15)
    return fail(MDE_ERR_INVALID, "gemm: operands must be 16-byte aligned");
  memset(&op->p, 0, sizeof(op->p));
  GemmParams& p = op->p;
  const int tok_skip = ep->token_skip > 0 ? ep->token_skip : 1;
  MDE_TRY(fill_epilogue(p, ep, ep->shuffle_s > 0 ? m * ep->shuffle_s * ep->shuffle_s : (ep->tokens > 0 ? m + m / ep->tokens * tok_skip : m)));
  p.M = static_cast<int>(m); p.N = n; p.K = k;
  if (m > 0x7fffffffLL) return fail(MDE_ERR_INVALID, "gemm: m too large");
  p.num_k_blocks = (k + 63) / 64;
  if (ep->d_head_w && n != 32) return fail(MDE_ERR_INVALID, "fused depth head needs n == 32");
  p.m_tiles = static_cast<int>((m + 127) / 128);
  pick_tiling(op, n, ep->d_head_w != nullptr, ep->gather_n > 0);
  p.conv = 0;
  if (ep->tokens > 0) {
    p.row_map = ROW_TOKENS; p.tokens = ep->tokens; p.tok_skip = tok_skip;
    if (m % ep->tokens) return fail(MDE_ERR_INVALID, "gemm: m not a multiple of tokens");
  } else if (ep->shuffle_s > 0) {
    p.row_map = ROW_SHUFFLE; p.shuffle_s = ep->shuffle_s; p.shuffle_cout = ep->shuffle_cout;
    p.H = ep->shuffle_h; p.W = ep->shuffle_w;
    if (n != ep->shuffle_s * ep->shuffle_s * ep->shuffle_cout || ep->shuffle_cout % 8 ||
        m % (static_cast<long long>(p.H) * p.W))
      return fail(MDE_ERR_INVALID, "gemm: inconsistent pixel-shuffle description");
  }
  op->precision = precision;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(m)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(lda) * 2};
    cuuint32_t box[2] = {64, 128};
    MDE_TRY(encode_map(&op->map_a, precision, d_a, 2, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(n)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(ldb) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(op->block_n / op->ctas)};   // each CTA of a pair stages half of the B tile
    MDE_TRY(encode_map(&op->map_b, precision, d_b, 2, dims, str, box));
  }
  MDE_TRY(maybe_tma_out(op, precision, ep, m, n));
  if (ep->gather_n > 0 && !p.gather_n) return fail(MDE_ERR_INVALID, "gemm: the fused gather needs a plain 16-bit output (bias / activation only) with N a multiple of the tile width");
  return pick_grid(op);
}

int make_conv_op(GemmOp* op, int precision, const void* d_in, int batch, int h, int w, int cin, const void* d_w,
                 int cout, const mde_epilogue* ep) {
  if (precision != MDE_FP16 && precision != MDE_BF16) return fail(MDE_ERR_INVALID, "precision must be MDE_FP16 or MDE_BF16");
  if (batch <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return fail(MDE_ERR_INVALID, "conv: empty problem");
  if (cin % 8 || cout % 8) return fail(MDE_ERR_INVALID, "conv: cin and cout must be multiples of 8");
  if (ep->tokens > 0 || ep->shuffle_s > 0) return fail(MDE_ERR_INVALID, "conv: row remaps are GEMM-only");
  memset(&op->p, 0, sizeof(op->p));
  GemmParams& p = op->p;
  MDE_TRY(fill_epilogue(p, ep, static_cast<long long>(batch) * h * w));
  const int cin_pad = (cin + 63) / 64 * 64;
  p.M = batch * h * w; p.N = cout; p.K = 9 * cin_pad;
  p.num_k_blocks = 9 * (cin_pad / 64);
  p.cin_blocks = cin_pad / 64;
  p.conv = 1; p.row_map = ROW_CONV; p.H = h; p.W = w;
  // spatial tile of <= 128 output pixels: maximise useful pixels per tile
  double best_eff = -1.0;
  for (int tw = 1; tw <= 128 && tw <= 256; ++tw) {
    const int th = 128 / tw;
    if (th < 1) break;
    if (th > 256) continue;
    const int tx = (w + tw - 1) / tw, ty = (h + th - 1) / th;
    const double eff = static_cast<double>(h) * w / (static_cast<double>(tx) * ty * 128.0);
    // prefer wide tiles on ties: longer contiguous runs per TMA box row
    if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && tw > p.tile_w)) {
      best_eff = eff; p.tile_w = tw; p.tile_h = th; p.tiles_x = tx; p.tiles_y = ty;
    }
  }
  if (ep->d_head_w && cout != 32) return fail(MDE_ERR_INVALID, "fused depth head needs cout == 32");
  p.m_tiles = batch * p.tiles_x * p.tiles_y;
  pick_tiling(op, cout, ep->d_head_w != nullptr);
  op->precision = precision;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                          static_cast<cuuint64_t>(batch)};
    cuuint64_t str[3] = {static_cast<cuuint64_t>(cin) * 2, static_cast<cuuint64_t>(w) * cin * 2,
                         static_cast<cuuint64_t>(h) * w * cin * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.tile_w), static_cast<cuuint32_t>(p.tile_h), 1};
    MDE_TRY(encode_map(&op->map_a, precision, d_in, 4, dims, str, box));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(9 * cin_pad), static_cast<cuuint64_t>(cout)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(9 * cin_pad) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(op->block_n / op->ctas)};
    MDE_TRY(encode_map(&op->map_b, precision, d_w, 2, dims, str, box));
  }
  memset(&op->map_out, 0, sizeof(op->map_out));
  return pick_grid(op);
}

// Launch with programmatic stream serialisation (see ptx.cuh `griddep_wait`): the kernel may be scheduled while its
// predecessor drains, so barrier initialisation, TMEM allocation and descriptor prefetch overlap the predecessor's tail.
// Only for kernels that call griddep_wait() before their first dependent access.  On for engine enqueues unless
// MDE_FLAG_NO_PDL is set or the process runs under MDE_PROFILE=1 (Nsight Compute 2025.2 does not list kernels launched
// with the attribute; numerics are identical either way).
static bool pdl_enabled() { return g_opts.pdl; }
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  int n = 0;
  if (cluster > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster; attrs[n].val.clusterDim.y = 1; attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attrs; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int BN, typename T, int kCtas>
static int launch_gemm_t(const GemmOp& op, cudaStream_t s) {
  auto kern = gemm_tcgen05_kernel<BN, T, kCtas>;
  using Cfg = GemmCfg<BN, kCtas>;
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(kern), Cfg::kSmemBytes, false));
  MDE_CUDA_TRY(launch_pdl(kern, dim3(op.grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, s, kCtas, op.map_a, op.map_b, op.map_out, op.gather, op.p));
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

template <typename T>
static int launch_gemm_bn(const GemmOp& op, cudaStream_t s) {
  if (op.ctas == 2) {
    switch (op.block_n) {
      case 256: return launch_gemm_t<256, T, 2>(op, s);
      case 128: return launch_gemm_t<128, T, 2>(op, s);
    }
    return fail(MDE_ERR_INVALID, "CTA pairs need BLOCK_N 128 or 256, not %d", op.block_n);
  }
  switch (op.block_n) {
    case 256: return launch_gemm_t<256, T, 1>(op, s);
    case 128: return launch_gemm_t<128, T, 1>(op, s);
    case 64: return launch_gemm_t<64, T, 1>(op, s);
    case 32: return launch_gemm_t<32, T, 1>(op, s);
  }
  return fail(MDE_ERR_INVALID, "unsupported BLOCK_N %d", op.block_n);
}

int launch_gemm(const GemmOp& op, cudaStream_t s) {
  return op.precision == MDE_BF16 ? launch_gemm_bn<__nv_bfloat16>(op, s) : launch_gemm_bn<__half>(op, s);
}

// ------------------------------------------------------------------------------------------- other launchers
template <typename T>
static int launch_attention_t(const void* d_qkv, void* d_out, int batch, int ntok, int heads, cudaStream_t s) {
  auto kern = attention_mma_kernel<T>;
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(kern), kAttnSmemBytes, false));
  AttnParams p;
  p.qkv = d_qkv; p.out = d_out; p.ntok = ntok; p.heads = heads; p.D = heads * 64;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  dim3 grid((ntok + kAttnBlockQ - 1) / kAttnBlockQ, heads, batch);
  kern<<<grid, 256, kAttnSmemBytes, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
int launch_attention_mma(int precision, const void* d_qkv, void* d_out, int batch, int ntok, int heads, cudaStream_t s) {
  if (batch <= 0 || ntok <= 0 || heads <= 0) return fail(MDE_ERR_INVALID, "attention: empty problem");
  if (batch > 65535 || heads > 65535) return fail(MDE_ERR_INVALID, "attention: batch/heads exceed grid limits");
  return precision == MDE_BF16 ? launch_attention_t<__nv_bfloat16>(d_qkv, d_out, batch, ntok, heads, s)
                               : launch_attention_t<__half>(d_qkv, d_out, batch, ntok, heads, s);
}

int make_attention_op_kv(AttnOp* op, int precision, const void* d_q, int ldq, const void* d_kv, int ldkv, int k_col0, int v_col0,
                         void* d_out, int batch, int ntok_q, int ntok_kv, int heads) {
  if (precision != MDE_FP16 && precision != MDE_BF16) return fail(MDE_ERR_INVALID, "precision must be MDE_FP16 or MDE_BF16");
  if (batch <= 0 || ntok_q <= 0 || ntok_kv <= 0 || heads <= 0) return fail(MDE_ERR_INVALID, "attention: empty problem");
  if (batch > 65535 || heads > 65535) return fail(MDE_ERR_INVALID, "attention: batch/heads exceed grid limits");
  if ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_kv) | reinterpret_cast<uintptr_t>(d_out)) & 15)
    return fail(MDE_ERR_INVALID, "attention: buffers must be 16-byte aligned");
  const int D = heads * 64;
  if (ldq < D || ldq % 8 || ldkv % 8 || k_col0 < 0 || v_col0 < 0 || k_col0 % 8 || v_col0 % 8 || k_col0 + D > ldkv || v_col0 + D > ldkv)
    return fail(MDE_ERR_INVALID, "attention: inconsistent pitches / column offsets");
  const long long rows_q = static_cast<long long>(batch) * ntok_q, rows_kv = static_cast<long long>(batch) * ntok_kv;
  if (rows_q > 0x7fffffffLL || rows_kv > 0x7fffffffLL) return fail(MDE_ERR_INVALID, "attention: too many rows");
  op->qkv = d_q; op->out = d_out; op->batch = batch; op->ntok = ntok_kv; op->heads = heads; op->precision = precision;
  op->ntok_q = ntok_q; op->k_col0 = k_col0; op->v_col0 = v_col0;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ldq), static_cast<cuuint64_t>(rows_q)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(ldq) * 2};
    cuuint32_t box[2] = {64, 128};
    MDE_TRY(encode_map(&op->map_qkv, precision, d_q, 2, dims, str, box));
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(ldkv), static_cast<cuuint64_t>(rows_kv)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(ldkv) * 2};
  cuuint32_t box128[2] = {64, 128};
  cuuint32_t box96[2] = {64, kAq3Keys};
  op->poly = 2;      // the measured optimum (tools/attn_sweep.py); engines override it from mde_engine_desc.attn_poly
  // Which kernel: one query tile per CTA, two CTAs per SM (attention_tc.cuh), or three query tiles per persistent CTA
  // (attention_q3.cuh).  At the batch-64 ViT-L shape the two take the same time; the persistent one reads each K / V tile
  // once for three query tiles (7 % less energy per launch, profiles/r02_energy_per_kernel.txt), which is what counts inside
  // the power-capped step.  Its work items are coarse, though (three query tiles over all keys, ~25 us at 1374 tokens), and a
  // launch takes whole rounds of them: measured at 1374 tokens and 16 heads (profiles/r02b_attention_small_batch.txt) the
  // one-tile kernel grows smoothly with the batch (48 / 56 / 65 / 108 / 188 us at batch 2 / 3 / 4 / 8 / 16) while the
  // persistent one steps at every multiple of 148 items (48 / 74 / 76 / 124 / 198 us).  So: the persistent kernel from eight
  // rounds on, where its last, partly filled round costs less than its shared K / V reads save.
  {
    const long long items = static_cast<long long>(batch) * heads * (((ntok_q + 127) / 128 + 2) / 3);
    op->kind = items >= 8LL * std::max(1, num_sms()) ? 1 : 0;
  }
  op->counters = nullptr;     // static item schedule unless the owner of the op gives it a work counter (engine.cu)
  MDE_TRY(encode_map(&op->map_kv96, precision, d_kv, 2, dims, str, box96));
  {
    const int D = heads * 64;
    cuuint64_t odims[3] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(ntok_q), static_cast<cuuint64_t>(batch)};
    cuuint64_t ostr[2] = {static_cast<cuuint64_t>(D) * 2, static_cast<cuuint64_t>(ntok_q) * D * 2};
    cuuint32_t obox[3] = {64, 32, 1};
    MDE_TRY(encode_map(&op->map_out3, precision, d_out, 3, odims, ostr, obox));
  }
  return encode_map(&op->map_kv128, precision, d_kv, 2, dims, str, box128);
}

int make_attention_op(AttnOp* op, int precision, const void* d_qkv, void* d_out, int batch, int ntok, int heads) {
  const int D = heads * 64;
  return make_attention_op_kv(op, precision, d_qkv, 3 * D, d_qkv, 3 * D, D, 2 * D, d_out, batch, ntok, ntok, heads);
}

template <typename T, int kPoly>
static int launch_attention_tc_t(const AttnOp& op, cudaStream_t s) {
  auto kern = attention_tc_kernel<T, kPoly>;
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(kern), kAtcSmemBytes, true));
  AttnParams p;
  p.qkv = op.qkv; p.out = op.out; p.ntok = op.ntok; p.heads = op.heads; p.D = op.heads * 64; p.trace = nullptr; p.batch = op.batch;
  p.ntok_q = op.ntok_q; p.k_col0 = op.k_col0; p.v_col0 = op.v_col0;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  dim3 grid((op.ntok_q + 127) / 128, op.heads, op.batch);
  MDE_CUDA_TRY(launch_pdl(kern, grid, dim3(kAtcThreads), kAtcSmemBytes, s, 1, op.map_qkv, op.map_kv128, p));
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
// kPoly of every 8 element pairs of the softmax take the FMA-pipe polynomial instead of the SFU (op.poly, 0..4; the
// engine's default is 2, the measured optimum).
template <typename T>
static int launch_attention_tc_p(const AttnOp& op, cudaStream_t s) {
  switch (op.poly) {
    case 0: return launch_attention_tc_t<T, 0>(op, s);
    case 1: return launch_attention_tc_t<T, 1>(op, s);
    case 2: return launch_attention_tc_t<T, 2>(op, s);
    case 3: return launch_attention_tc_t<T, 3>(op, s);
    case 4: return launch_attention_tc_t<T, 4>(op, s);
  }
  return fail(MDE_ERR_INVALID, "attention: the polynomial share is 0..4 eighths, not %d", op.poly);
}
template <typename T>
static int launch_attention_trace_t(const AttnOp& op, long long* d_trace, cudaStream_t s) {
  auto kern = attention_tc_kernel<T, 2, true>;
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(kern), kAtcSmemBytes, true));
  AttnParams p;
  p.qkv = op.qkv; p.out = op.out; p.ntok = op.ntok; p.heads = op.heads; p.D = op.heads * 64; p.trace = d_trace; p.batch = op.batch;
  p.ntok_q = op.ntok_q; p.k_col0 = op.k_col0; p.v_col0 = op.v_col0;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  dim3 grid((op.ntok_q + 127) / 128, op.heads, op.batch);
  kern<<<grid, kAtcThreads, kAtcSmemBytes, s>>>(op.map_qkv, op.map_kv128, p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

template <typename T, int kPoly>
static int launch_attention_q3_t(const AttnOp& op, cudaStream_t s) {
  auto kern = attention_q3_kernel<T, kPoly>;
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(kern), kAq3SmemBytes, true));
  AttnParams p;
  p.qkv = op.qkv; p.out = op.out; p.ntok = op.ntok; p.heads = op.heads; p.D = op.heads * 64; p.trace = nullptr;
  p.ntok_q = op.ntok_q; p.k_col0 = op.k_col0; p.v_col0 = op.v_col0; p.batch = op.batch;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  const int q_tiles = (op.ntok_q + 127) / 128;
  const long long items = static_cast<long long>(op.batch) * op.heads * ((q_tiles + 2) / 3);
  if (items > 0x3fffffffLL) return fail(MDE_ERR_INVALID, "attention: too many work items");
  dim3 grid(static_cast<unsigned>(std::min<long long>(items, num_sms())));
  MDE_CUDA_TRY(launch_pdl(kern, grid, dim3(kAq3Threads), kAq3SmemBytes, s, 1, op.map_qkv, op.map_kv96, op.map_out3, p, op.counters));
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
template <typename T>
static int launch_attention_q3_trace_t(const AttnOp& op, long long* d_trace, cudaStream_t s) {
  auto kern = attention_q3_kernel<T, 2, true>;
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(kern), kAq3SmemBytes, true));
  AttnParams p;
  p.qkv = op.qkv; p.out = op.out; p.ntok = op.ntok; p.heads = op.heads; p.D = op.heads * 64; p.trace = d_trace; p.batch = op.batch;
  p.ntok_q = op.ntok_q; p.k_col0 = op.k_col0; p.v_col0 = op.v_col0;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  const int q_tiles = (op.ntok_q + 127) / 128;
  const long long items = static_cast<long long>(op.batch) * op.heads * ((q_tiles + 2) / 3);
  kern<<<static_cast<unsigned>(std::min<long long>(items, num_sms())), kAq3Threads, kAq3SmemBytes, s>>>(op.map_qkv, op.map_kv96, op.map_out3, p, op.counters);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
template <typename T>
static int launch_attention_q3_p(const AttnOp& op, cudaStream_t s) {
  switch (op.poly) {
    case 0: return launch_attention_q3_t<T, 0>(op, s);
    case 1: return launch_attention_q3_t<T, 1>(op, s);
    case 2: return launch_attention_q3_t<T, 2>(op, s);
    case 3: return launch_attention_q3_t<T, 3>(op, s);
    case 4: return launch_attention_q3_t<T, 4>(op, s);
  }
  return fail(MDE_ERR_INVALID, "attention: the polynomial share is 0..4 eighths, not %d", op.poly);
}

int launch_attention_op(const AttnOp& op, cudaStream_t s) {
  if (op.kind == 1) return op.precision == MDE_BF16 ? launch_attention_q3_p<__nv_bfloat16>(op, s) : launch_attention_q3_p<__half>(op, s);
  return op.precision == MDE_BF16 ? launch_attention_tc_p<__nv_bfloat16>(op, s) : launch_attention_tc_p<__half>(op, s);
}

template <typename T, bool kTap>
static int launch_layernorm_t(const LayerNormParams& p, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((p.rows + 7) / 8);
  switch (p.D) {
    case 384: MDE_CUDA_TRY(launch_pdl(layernorm_kernel<T, 384, kTap>, dim3(grid), dim3(256), 0, s, 1, p)); break;
    case 768: MDE_CUDA_TRY(launch_pdl(layernorm_kernel<T, 768, kTap>, dim3(grid), dim3(256), 0, s, 1, p)); break;
    case 1024: MDE_CUDA_TRY(launch_pdl(layernorm_kernel<T, 1024, kTap>, dim3(grid), dim3(256), 0, s, 1, p)); break;
    case 128: MDE_CUDA_TRY(launch_pdl(layernorm_kernel<T, 128, kTap>, dim3(grid), dim3(256), 0, s, 1, p)); break;
    case 1536: MDE_CUDA_TRY(launch_pdl(layernorm_kernel<T, 1536, kTap>, dim3(grid), dim3(256), 0, s, 1, p)); break;
    case 2048: MDE_CUDA_TRY(launch_pdl(layernorm_kernel<T, 2048, kTap>, dim3(grid), dim3(256), 0, s, 1, p)); break;
    default: return fail(MDE_ERR_INVALID, "layernorm: unsupported width %d", p.D);
  }
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}
int launch_layernorm(int precision, const float* d_x, const float* d_w, const float* d_b, void* d_out, long long rows,
                     int dim, float eps, int drop_cls, int ntok, cudaStream_t s, int identity, int n_dst, void* const* dst,
                     long long dst_row0) {
  if (rows <= 0) return fail(MDE_ERR_INVALID, "layernorm: no rows");
  if (drop_cls < 0 || (drop_cls && (ntok <= drop_cls || rows % ntok))) return fail(MDE_ERR_INVALID, "layernorm: rows not a multiple of ntok, or nothing left after dropping %d tokens", drop_cls);
  if (n_dst < 0 || n_dst > 8 || (n_dst > 0 && !dst)) return fail(MDE_ERR_INVALID, "layernorm: at most 8 gather destinations");
  LayerNormParams p;
  p.x = d_x; p.w = d_w; p.b = d_b; p.out = d_out; p.rows = rows; p.D = dim; p.eps = eps; p.drop_cls = drop_cls; p.ntok = ntok;
  p.identity = identity; p.n_dst = n_dst; p.dst_row0 = dst_row0;
  for (int i = 0; i < 8; ++i) p.dst[i] = i < n_dst ? dst[i] : nullptr;
  if (identity || n_dst > 0)
    return precision == MDE_BF16 ? launch_layernorm_t<__nv_bfloat16, true>(p, s) : launch_layernorm_t<__half, true>(p, s);
  return precision == MDE_BF16 ? launch_layernorm_t<__nv_bfloat16, false>(p, s) : launch_layernorm_t<__half, false>(p, s);
}

static unsigned grid_for(long long total, int per_block) {
  const long long blocks = (total + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(std::max(1, num_sms())) * 16;
  return static_cast<unsigned>(std::max(1LL, std::min(blocks, cap)));
}

int launch_bilinear(int precision, const void* d_in, void* d_out, int batch, int hi, int wi, int ho, int wo, int c,
                    cudaStream_t s, const float* d_addend) {
  if (c % 8) return fail(MDE_ERR_INVALID, "bilinear: channels must be a multiple of 8");
  if (batch <= 0 || hi <= 0 || wi <= 0 || ho <= 0 || wo <= 0) return fail(MDE_ERR_INVALID, "bilinear: empty problem");
  BilinearParams p;
  p.in = d_in; p.out = d_out; p.B = batch; p.Hi = hi; p.Wi = wi; p.Ho = ho; p.Wo = wo; p.C = c; p.addend = d_addend;
  if (d_addend && (reinterpret_cast<uintptr_t>(d_addend) & 15)) return fail(MDE_ERR_INVALID, "bilinear: the addend must be 16-byte aligned");
  p.sy = ho > 1 ? static_cast<float>(hi - 1) / static_cast<float>(ho - 1) : 0.f;
  p.sx = wo > 1 ? static_cast<float>(wi - 1) / static_cast<float>(wo - 1) : 0.f;
  if (batch > 65535) return fail(MDE_ERR_INVALID, "bilinear: batch exceeds grid limits");
  dim3 grid(ho, batch);
  if (precision == MDE_BF16) bilinear_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  else bilinear_nhwc_kernel<__half><<<grid, 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_resize_depth(const float* d_in, int batch, int hi, int wi, float* d_out, int ho, int wo, float lo, float hi_clamp,
                        cudaStream_t s) {
  if (batch <= 0 || hi <= 0 || wi <= 0 || ho <= 0 || wo <= 0) return fail(MDE_ERR_INVALID, "resize_depth: empty problem");
  if (batch > 65535 || ho > 65535) return fail(MDE_ERR_INVALID, "resize_depth: batch / height exceed grid limits");
  ResizeDepthParams p;
  p.in = d_in; p.out = d_out; p.B = batch; p.Hi = hi; p.Wi = wi; p.Ho = ho; p.Wo = wo; p.lo = lo; p.hi = hi_clamp;
  p.sy = ho > 1 ? static_cast<float>(hi - 1) / static_cast<float>(ho - 1) : 0.f;
  p.sx = wo > 1 ? static_cast<float>(wi - 1) / static_cast<float>(wo - 1) : 0.f;
  dim3 grid((wo + 255) / 256, ho, batch);
  resize_depth_kernel<<<grid, 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_qknorm_rope(int precision, void* d_qkv, long long rows, int heads, const float* d_qw, const float* d_qb, const float* d_kw,
                       const float* d_kb, float eps, const int* d_pos, const float* d_cos_sin, int max_pos, int gather_n,
                       void* const* d_gather, int gather_ld, cudaStream_t s) {
  if (rows <= 0 || heads <= 0) return fail(MDE_ERR_INVALID, "qknorm_rope: empty problem");
  if (rows > 8LL * 0x7fffffffLL) return fail(MDE_ERR_INVALID, "qknorm_rope: too many rows");
  if (d_pos && (!d_cos_sin || max_pos < 1)) return fail(MDE_ERR_INVALID, "qknorm_rope: positions need the cos / sin table");
  if (gather_n < 0 || gather_n > 8 || (gather_n > 0 && (!d_gather || gather_ld < 2 * heads * 64 || gather_ld % 8)))
    return fail(MDE_ERR_INVALID, "qknorm_rope: at most 8 gather destinations with a pitch >= 2 * D, a multiple of 8");
  if (reinterpret_cast<uintptr_t>(d_qkv) & 15) return fail(MDE_ERR_INVALID, "qknorm_rope: rows must be 16-byte aligned");
  QkNormRopeParams p;
  memset(&p, 0, sizeof(p));
  p.qkv = d_qkv; p.qw = d_qw; p.qb = d_qb; p.kw = d_kw; p.kb = d_kb; p.pos = d_pos; p.cos_sin = d_cos_sin; p.rows = rows;
  p.heads = heads; p.max_pos = max_pos; p.gather_n = gather_n; p.gather_ld = gather_ld; p.eps = eps;
  for (int r = 0; r < gather_n; ++r) {
    if (!d_gather[r] || (reinterpret_cast<uintptr_t>(d_gather[r]) & 15)) return fail(MDE_ERR_INVALID, "qknorm_rope: gather destination %d is null or misaligned", r);
    p.gather[r] = d_gather[r];
  }
  const int sms = num_sms();
  if (sms <= 0) return fail(MDE_ERR_CUDA, "no CUDA device");
  const dim3 grid(static_cast<unsigned>(std::min<long long>((rows + 7) / 8, 2LL * sms)));   // persistent warps, two CTAs per SM
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(qknorm_rope_kernel<__nv_bfloat16>), kRopeSmemBytes, false));
  MDE_TRY(ensure_func_attrs(reinterpret_cast<const void*>(qknorm_rope_kernel<__half>), kRopeSmemBytes, false));
  if (precision == MDE_BF16) MDE_CUDA_TRY(launch_pdl(qknorm_rope_kernel<__nv_bfloat16>, grid, dim3(256), kRopeSmemBytes, s, 1, p));
  else MDE_CUDA_TRY(launch_pdl(qknorm_rope_kernel<__half>, grid, dim3(256), kRopeSmemBytes, s, 1, p));
  return MDE_OK;
}

int launch_resize_crops(const void* d_src, int src_u8, int swap_rb, int src_h, int src_w, const mde_crop* crops, int n_crops,
                        int out_h, int out_w, const float* mean3, const float* std3, float* d_out, cudaStream_t s) {
  if (src_h < 1 || src_w < 1 || out_h < 1 || out_w < 1) return fail(MDE_ERR_INVALID, "resize_crops: empty problem");
  if (n_crops < 1 || n_crops > 64) return fail(MDE_ERR_INVALID, "resize_crops: 1..64 crops per launch, not %d", n_crops);
  if (out_h > 65535 || n_crops * 3 > 65535) return fail(MDE_ERR_INVALID, "resize_crops: output exceeds grid limits");
  if ((mean3 == nullptr) != (std3 == nullptr)) return fail(MDE_ERR_INVALID, "resize_crops: mean and std go together (both NULL: no normalisation)");
  if (mean3 && !src_u8) return fail(MDE_ERR_INVALID, "resize_crops: normalisation applies to the uint8 source only");
  ResizeCropsParams p;
  memset(&p, 0, sizeof(p));
  p.src = d_src; p.out = d_out; p.src_u8 = src_u8 ? 1 : 0; p.swap_rb = swap_rb ? 1 : 0; p.normalise = mean3 ? 1 : 0;
  p.src_h = src_h; p.src_w = src_w; p.out_h = out_h; p.out_w = out_w; p.n_crops = n_crops;
  for (int i = 0; i < 3; ++i) { p.mean[i] = mean3 ? mean3[i] : 0.f; p.std[i] = std3 ? std3[i] : 1.f; }
  for (int i = 0; i < n_crops; ++i) {
    const mde_crop& c = crops[i];
    if (c.level_h < 1 || c.level_w < 1 || c.y0 < 0 || c.x0 < 0 || c.y0 + out_h > c.level_h || c.x0 + out_w > c.level_w)
      return fail(MDE_ERR_INVALID, "resize_crops: crop %d (%d,%d)+(%dx%d) leaves its %dx%d level", i, c.y0, c.x0, out_h, out_w, c.level_h, c.level_w);
    p.crops[i].level_h = c.level_h; p.crops[i].level_w = c.level_w; p.crops[i].y0 = c.y0; p.crops[i].x0 = c.x0;
  }
  dim3 grid((out_w + 255) / 256, out_h, n_crops * 3);
  resize_crops_kernel<<<grid, 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_depth_pro_post(const float* d_inv, const float* d_fov, int h, int w, int src_h, int src_w, float* d_depth, float* d_f_px,
                          cudaStream_t s) {
  if (h < 1 || w < 1 || src_h < 1 || src_w < 1) return fail(MDE_ERR_INVALID, "depth_pro_post: empty problem");
  if (src_h > 65535) return fail(MDE_ERR_INVALID, "depth_pro_post: height exceeds grid limits");
  DepthProPostParams p;
  p.inv = d_inv; p.fov_deg = d_fov; p.depth = d_depth; p.f_px = d_f_px; p.h = h; p.w = w; p.pitch = w; p.src_h = src_h; p.src_w = src_w;
  p.mul = 1.f; p.lo = 1e-4f; p.hi = 1e4f; p.reciprocal = 1; p.nan_below = 0;
  dim3 grid((src_w + 255) / 256, src_h, 1);
  depth_pro_post_kernel<<<grid, 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_resize_depth_halfpixel(const float* d_in, int pitch, int h, int w, float* d_out, int out_h, int out_w, float mul, float lo,
                                  float hi, cudaStream_t s, int nan_below = 0) {
  if (h < 1 || w < 1 || out_h < 1 || out_w < 1 || pitch < w) return fail(MDE_ERR_INVALID, "resize_depth_halfpixel: empty problem or pitch < width");
  if (out_h > 65535) return fail(MDE_ERR_INVALID, "resize_depth_halfpixel: height exceeds grid limits");
  DepthProPostParams p;
  p.inv = d_in; p.fov_deg = nullptr; p.depth = d_out; p.f_px = nullptr; p.h = h; p.w = w; p.pitch = pitch; p.src_h = out_h; p.src_w = out_w;
  p.mul = mul; p.lo = lo; p.hi = hi; p.reciprocal = 0; p.nan_below = nan_below ? 1 : 0;
  dim3 grid((out_w + 255) / 256, out_h, 1);
  depth_pro_post_kernel<<<grid, 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_merge_patches(int precision, const void* d_in, int per_side, int grid, int pad, int dim, void* d_out, cudaStream_t s) {
  if (per_side < 1 || grid < 1 || dim < 8 || dim % 8) return fail(MDE_ERR_INVALID, "merge_patches: bad geometry");
  if (per_side == 1) pad = 0;
  if (pad < 0 || pad > grid / 4) return fail(MDE_ERR_INVALID, "merge_patches: padding must be within [0, grid/4]");
  if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) return fail(MDE_ERR_INVALID, "merge_patches: buffers must be 16-byte aligned");
  MergeParams p;
  p.in = d_in; p.out = d_out; p.per_side = per_side; p.grid = grid; p.pad = pad; p.D = dim;
  p.S = per_side * grid - 2 * pad * (per_side - 1);
  const long long total = static_cast<long long>(p.S) * p.S * (dim / 8);
  if (precision == MDE_BF16) merge_patches_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(p);
  else merge_patches_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_im2col_s2(int precision, const void* d_in, void* d_out, int batch, int h, int w, int c, cudaStream_t s) {
  if (c % 8) return fail(MDE_ERR_INVALID, "im2col_s2: channels must be a multiple of 8");
  Im2colS2Params p;
  p.in = d_in; p.out = d_out; p.B = batch; p.H = h; p.W = w; p.C = c;
  p.Ho = (h - 1) / 2 + 1; p.Wo = (w - 1) / 2 + 1;
  const long long total = static_cast<long long>(batch) * p.Ho * p.Wo * 9 * (c / 8);
  if (precision == MDE_BF16) im2col_s2_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, s>>>(p);
  else im2col_s2_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

static int check_patch_geometry(int h, int w, int patch, int kpad) {
  if (patch <= 0 || h % patch || w % patch) return fail(MDE_ERR_INVALID, "input size %dx%d is not a multiple of patch %d", h, w, patch);
  if (kpad % 64 || kpad < 3 * patch * patch) return fail(MDE_ERR_INVALID, "kpad %d must be a multiple of 64 and >= %d", kpad, 3 * patch * patch);
  return MDE_OK;
}

int launch_im2col_f32(int precision, const float* d_nchw, int batch, int h, int w, int patch, int kpad, void* d_cols,
                      cudaStream_t s, const double* mean3, const double* std3) {
  MDE_TRY(check_patch_geometry(h, w, patch, kpad));
  Im2colParams p;
  p.nchw = d_nchw; p.cols = d_cols; p.H = h; p.W = w; p.patch = patch; p.kpad = kpad;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3 ? static_cast<float>(mean3[c]) : 0.f;
    p.inv_std[c] = std3 ? static_cast<float>(1.0 / std3[c]) : 1.f;
  }
  const int smem = (w / patch) * kpad * 2;
  dim3 grid(h / patch, batch);
  if (precision == MDE_BF16) {
    MDE_CUDA_TRY(cudaFuncSetAttribute(im2col_f32_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    im2col_f32_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(p);
  } else {
    MDE_CUDA_TRY(cudaFuncSetAttribute(im2col_f32_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    im2col_f32_kernel<__half><<<grid, 256, smem, s>>>(p);
  }
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

void build_norm_lut(const double* mean3, const double* std3, float* lut768, bool scale_f32) {
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      volatile float xf = static_cast<float>(v) / 255.0f;   // depth_anything_ac divides in float32 (core/preprocess.py:294-305)
      volatile double x = scale_f32 ? static_cast<double>(xf) : static_cast<double>(v) / 255.0;   // volatile: no fused/reassociated evaluation
      volatile double y = x - mean3[c];
      volatile double z = y / std3[c];
      lut768[c * 256 + v] = static_cast<float>(z);
    }
}

void build_raw_lut(float* lut768) {
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) lut768[c * 256 + v] = static_cast<float>(v);
}

int launch_preprocess_u8(int precision, const uint8_t* d_src, long long src_batch_stride, int batch, int src_h,
                         int src_w, int dst_h, int dst_w, int patch, int kpad, int swap_rb, const float* d_lut,
                         void* d_cols, float* d_nchw, cudaStream_t s, int keep_ratio_pad, const double* pad_rgb) {
  MDE_TRY(check_patch_geometry(dst_h, dst_w, patch, kpad));
  if (patch > 16) return fail(MDE_ERR_INVALID, "preprocess_u8: patch sizes up to 16, not %d", patch);
  if (src_h < 1 || src_w < 1) return fail(MDE_ERR_INVALID, "preprocess: empty source image");
  PreprocParams p;
  p.src = d_src; p.src_batch_stride = src_batch_stride; p.src_h = src_h; p.src_w = src_w;
  p.dst_h = dst_h; p.dst_w = dst_w; p.patch = patch; p.kpad = kpad; p.swap_rb = swap_rb; p.lut = d_lut;
  p.cols = d_cols; p.nchw = d_nchw;
  p.inner_h = dst_h; p.inner_w = dst_w; p.pad_top = 0; p.pad_left = 0;
  p.pad_src[0] = p.pad_src[1] = p.pad_src[2] = p.pad_src[3] = 0;
  if (keep_ratio_pad) {
    // core/preprocess.py:191-219 `resize_pad`, rounding='trunc', center=True -- the same double arithmetic as Python's
    if (!pad_rgb) return fail(MDE_ERR_INVALID, "preprocess: the pad colour is required");
    const double sh = static_cast<double>(dst_h) / static_cast<double>(src_h), sw = static_cast<double>(dst_w) / static_cast<double>(src_w);
    const double scale = sh < sw ? sh : sw;
    p.inner_h = static_cast<int>(static_cast<double>(src_h) * scale);
    p.inner_w = static_cast<int>(static_cast<double>(src_w) * scale);
    if (p.inner_h < 1 || p.inner_w < 1 || p.inner_h > dst_h || p.inner_w > dst_w) return fail(MDE_ERR_INVALID, "preprocess: degenerate keep-ratio geometry");
    p.pad_top = (dst_h - p.inner_h) / 2;
    p.pad_left = (dst_w - p.inner_w) / 2;
    for (int c = 0; c < 3; ++c) {
      // cv2.copyMakeBorder saturate-casts the value: round half to even, clamp to 0..255; stored in SOURCE channel order
      double v = nearbyint(pad_rgb[c]);
      v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
      p.pad_src[swap_rb ? 2 - c : c] = static_cast<uint8_t>(v);
    }
  }
  p.exact2x = (src_w == 2 * p.inner_w && src_h == 2 * p.inner_h) ? 1 : 0;
  const int smem = ((dst_w * 8 + 15) & ~15) + (d_cols ? (dst_w / patch) * kpad * 2 : 0);
  dim3 grid(dst_h / patch, batch);
  if (precision == MDE_BF16) {
    MDE_CUDA_TRY(cudaFuncSetAttribute(preprocess_u8_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    preprocess_u8_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(p);
  } else {
    MDE_CUDA_TRY(cudaFuncSetAttribute(preprocess_u8_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    preprocess_u8_kernel<__half><<<grid, 256, smem, s>>>(p);
  }
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_upconv_head(int precision, const void* d_z, int ldz, int batch, int hs, int ws, int ho, int wo,
                       const float* d_bias, const float* d_head_w, float head_b, float head_scale, float* d_out,
                       cudaStream_t s, int head_exp) {
  if (batch <= 0 || hs <= 0 || ws <= 0 || ho <= 0 || wo <= 0) return fail(MDE_ERR_INVALID, "upconv_head: empty problem");
  if (ldz < kUpZc || ldz % 8 || (reinterpret_cast<uintptr_t>(d_z) & 15)) return fail(MDE_ERR_INVALID, "upconv_head: z needs >= 288 channels, a pitch that is a multiple of 8 and 16-byte alignment");
  if (batch > 65535) return fail(MDE_ERR_INVALID, "upconv_head: batch exceeds grid limits");
  UpconvHeadParams p;
  p.z = d_z; p.out = d_out; p.bias = d_bias; p.head_w = d_head_w; p.head_b = head_b;
  p.head_scale = head_scale > 0.f ? head_scale : -1.f;
  p.head_exp = head_exp ? 1 : 0;
  p.B = batch; p.Hs = hs; p.Ws = ws; p.Ho = ho; p.Wo = wo; p.ldz = ldz;
  p.sy = ho > 1 ? static_cast<float>(hs - 1) / static_cast<float>(ho - 1) : 0.f;
  p.sx = wo > 1 ? static_cast<float>(ws - 1) / static_cast<float>(wo - 1) : 0.f;
  // staged window: the largest footprint over all tile origins, with exactly the kernel's arithmetic
  auto extent = [](int out_n, int src_n, float sc) {
    int best = 1;
    for (int o0 = 0; o0 < out_n; o0 += kUpTile) {
      const int lo = std::max(o0 - 1, 0), hi = std::min(o0 + kUpTile, out_n - 1);
      const int s_lo = std::min(static_cast<int>(lo * sc), src_n - 1);
      const int s_hi = std::min(std::min(static_cast<int>(hi * sc), src_n - 1) + 1, src_n - 1);
      best = std::max(best, s_hi - s_lo + 1);
    }
    return best;
  };
  p.fh = extent(ho, hs, p.sy);
  p.fw = extent(wo, ws, p.sx);
  const int smem = p.fh * p.fw * kUpPixBytes;
  if (smem > 113 * 1024) return fail(MDE_ERR_INVALID, "upconv_head: scale %dx%d -> %dx%d needs %d bytes of shared memory per tile", hs, ws, ho, wo, smem);
  dim3 grid((wo + kUpTile - 1) / kUpTile, (ho + kUpTile - 1) / kUpTile, batch);
  if (precision == MDE_BF16) {
    MDE_CUDA_TRY(cudaFuncSetAttribute(upconv_head_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    upconv_head_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(p);
  } else {
    MDE_CUDA_TRY(cudaFuncSetAttribute(upconv_head_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    upconv_head_kernel<__half><<<grid, 256, smem, s>>>(p);
  }
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int launch_cls_row(float* d_x, const float* d_cls, const float* d_pos, const float* d_reg, int n_reg, int batch, int ntok, int dim,
                   cudaStream_t s) {
  cls_row_kernel<<<batch, 256, 0, s>>>(d_x, d_cls, d_pos, d_reg, n_reg, ntok, dim);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

}  // namespace mde

// =============================================================================================== C ABI
using namespace mde;

extern "C" {

const char* mde_last_error(void) { return mde::last_error_cstr(); }
int mde_abi_version(void) { return MDE_ABI_VERSION; }

int mde_k_gemm(int32_t precision, const void* d_a, int64_t m, int32_t k, int32_t lda, const void* d_b, int32_t n,
               int32_t ldb, const mde_epilogue* ep, void* stream) {
  clear_error();
  if (!ep) return fail(MDE_ERR_INVALID, "epilogue description is required");
  GemmOp op;
  MDE_TRY(make_gemm_op(&op, precision, d_a, m, k, lda, d_b, n, ldb, ep));
  return launch_gemm(op, static_cast<cudaStream_t>(stream));
}

int mde_k_gemm_tiled(int32_t precision, const void* d_a, int64_t m, int32_t k, int32_t lda, const void* d_b, int32_t n,
                     int32_t ldb, const mde_epilogue* ep, int32_t block_n, int32_t ctas, int32_t splits, void* stream) {
  clear_error();
  if (!ep) return fail(MDE_ERR_INVALID, "epilogue description is required");
  if ((block_n != 32 && block_n != 64 && block_n != 128 && block_n != 256) || ctas < 1 || ctas > 2 || splits < 1 || splits > 16)
    return fail(MDE_ERR_INVALID, "gemm_tiled: block_n 32/64/128/256, ctas 1/2, splits 1..16");
  GemmOp op;
  g_tiling.block_n = block_n; g_tiling.ctas = ctas; g_tiling.splits = splits;
  const int rc = make_gemm_op(&op, precision, d_a, m, k, lda, d_b, n, ldb, ep);
  g_tiling = TilingOverride{};
  MDE_TRY(rc);
  return launch_gemm(op, static_cast<cudaStream_t>(stream));
}

int mde_k_conv3x3(int32_t precision, const void* d_in, int32_t batch, int32_t h, int32_t w, int32_t cin,
                  const void* d_w, int32_t cout, const mde_epilogue* ep, void* stream) {
  clear_error();
  if (!ep) return fail(MDE_ERR_INVALID, "epilogue description is required");
  GemmOp op;
  MDE_TRY(make_conv_op(&op, precision, d_in, batch, h, w, cin, d_w, cout, ep));
  return launch_gemm(op, static_cast<cudaStream_t>(stream));
}

int mde_k_attention(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                    void* stream) {
  clear_error();
  AttnOp op;
  MDE_TRY(make_attention_op(&op, precision, d_qkv, d_out, batch, ntok, heads));
  return launch_attention_op(op, static_cast<cudaStream_t>(stream));
}

int mde_k_attention_poly(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                         int32_t poly_eighths, void* stream) {
  clear_error();
  AttnOp op;
  MDE_TRY(make_attention_op(&op, precision, d_qkv, d_out, batch, ntok, heads));
  if (poly_eighths >= 0) op.poly = poly_eighths;
  op.kind = 0;
  return launch_attention_op(op, static_cast<cudaStream_t>(stream));
}

int mde_k_attention_q3(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                       int32_t poly_eighths, void* stream) {
  clear_error();
  AttnOp op;
  MDE_TRY(make_attention_op(&op, precision, d_qkv, d_out, batch, ntok, heads));
  if (poly_eighths >= 0) op.poly = poly_eighths;
  op.kind = 1;
  return launch_attention_op(op, static_cast<cudaStream_t>(stream));
}

int mde_k_attention_kv(int32_t precision, const void* d_q, int32_t ldq, const void* d_kv, int32_t ldkv, int32_t k_col0,
                       int32_t v_col0, void* d_out, int32_t batch, int32_t ntok_q, int32_t ntok_kv, int32_t heads, void* stream) {
  clear_error();
  AttnOp op;
  MDE_TRY(make_attention_op_kv(&op, precision, d_q, ldq, d_kv, ldkv, k_col0, v_col0, d_out, batch, ntok_q, ntok_kv, heads));
  return launch_attention_op(op, static_cast<cudaStream_t>(stream));
}

int mde_k_attention_trace(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                          int64_t* d_trace, void* stream) {
  clear_error();
  if (!d_trace) return fail(MDE_ERR_INVALID, "attention_trace: the trace buffer is required");
  AttnOp op;
  const bool q3 = precision >= 16;   // profiling aid: precision + 16 traces the three-query-tile kernel ([CTAs][16 rows][256 slots], attention_q3.cuh)
  if (q3) precision -= 16;
  MDE_TRY(make_attention_op(&op, precision, d_qkv, d_out, batch, ntok, heads));
  if (q3)
    return precision == MDE_BF16 ? launch_attention_q3_trace_t<__nv_bfloat16>(op, reinterpret_cast<long long*>(d_trace), static_cast<cudaStream_t>(stream))
                                      : launch_attention_q3_trace_t<__half>(op, reinterpret_cast<long long*>(d_trace), static_cast<cudaStream_t>(stream));
  return precision == MDE_BF16 ? launch_attention_trace_t<__nv_bfloat16>(op, reinterpret_cast<long long*>(d_trace), static_cast<cudaStream_t>(stream))
                               : launch_attention_trace_t<__half>(op, reinterpret_cast<long long*>(d_trace), static_cast<cudaStream_t>(stream));
}

int mde_k_attention_mma(int32_t precision, const void* d_qkv, void* d_out, int32_t batch, int32_t ntok, int32_t heads,
                        void* stream) {
  clear_error();
  return launch_attention_mma(precision, d_qkv, d_out, batch, ntok, heads, static_cast<cudaStream_t>(stream));
}

int mde_k_layernorm(int32_t precision, const float* d_x, const float* d_w, const float* d_b, void* d_out,
                    int64_t rows, int32_t dim, float eps, int32_t drop_cls, int32_t ntok, void* stream) {
  clear_error();
  return launch_layernorm(precision, d_x, d_w, d_b, d_out, rows, dim, eps, drop_cls, ntok,
                          static_cast<cudaStream_t>(stream));
}

int mde_k_bilinear(int32_t precision, const void* d_in, void* d_out, int32_t batch, int32_t hi, int32_t wi, int32_t ho,
                   int32_t wo, int32_t c, void* stream) {
  clear_error();
  return launch_bilinear(precision, d_in, d_out, batch, hi, wi, ho, wo, c, static_cast<cudaStream_t>(stream));
}

int mde_k_bilinear_add(int32_t precision, const void* d_in, void* d_out, int32_t batch, int32_t hi, int32_t wi, int32_t ho,
                       int32_t wo, int32_t c, const float* d_addend, void* stream) {
  clear_error();
  if (!d_addend) return fail(MDE_ERR_INVALID, "bilinear_add: the addend is required");
  return launch_bilinear(precision, d_in, d_out, batch, hi, wi, ho, wo, c, static_cast<cudaStream_t>(stream), d_addend);
}

int mde_k_assemble_tokens(int32_t precision, const void* d_patch, const float* d_special, int32_t frames, int32_t tokens,
                          int32_t n_special, int32_t dim, int32_t first_frame, float* d_out, void* stream) {
  clear_error();
  if (precision != MDE_FP16 && precision != MDE_BF16) return fail(MDE_ERR_INVALID, "precision must be MDE_FP16 or MDE_BF16");
  if (!d_patch || !d_special || !d_out) return fail(MDE_ERR_INVALID, "assemble_tokens: null pointer");
  if (frames < 1 || tokens < 1 || n_special < 1 || dim < 8 || dim % 8 || first_frame < 0) return fail(MDE_ERR_INVALID, "assemble_tokens: bad geometry");
  if ((reinterpret_cast<uintptr_t>(d_patch) | reinterpret_cast<uintptr_t>(d_special) | reinterpret_cast<uintptr_t>(d_out)) & 15)
    return fail(MDE_ERR_INVALID, "assemble_tokens: buffers must be 16-byte aligned");
  AssembleTokensParams p;
  p.patch = d_patch; p.special = d_special; p.out = d_out; p.frames = frames; p.tokens = tokens; p.n_special = n_special; p.D = dim;
  p.first_frame = first_frame;
  const unsigned grid = static_cast<unsigned>(frames) * static_cast<unsigned>(n_special + tokens);
  if (precision == MDE_BF16) assemble_tokens_kernel<__nv_bfloat16><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else assemble_tokens_kernel<__half><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_im2col_s2(int32_t precision, const void* d_in, void* d_out, int32_t batch, int32_t h, int32_t w, int32_t c,
                    void* stream) {
  clear_error();
  return launch_im2col_s2(precision, d_in, d_out, batch, h, w, c, static_cast<cudaStream_t>(stream));
}

int mde_k_peer_signal(void* const* d_flags_of_every_rank, int32_t n_ranks, int32_t rank, uint32_t epoch, void* stream) {
  clear_error();
  if (!d_flags_of_every_rank || n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks) return fail(MDE_ERR_INVALID, "peer_signal: 1..8 ranks, 0 <= rank < n_ranks");
  PeerFlagsParams p;
  for (int r = 0; r < 8; ++r) p.flags[r] = r < n_ranks ? static_cast<unsigned int*>(d_flags_of_every_rank[r]) : nullptr;
  for (int r = 0; r < n_ranks; ++r)
    if (!p.flags[r]) return fail(MDE_ERR_INVALID, "peer_signal: flag array of rank %d is null", r);
  p.n_ranks = n_ranks; p.rank = rank; p.epoch = epoch;
  peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_peer_wait(void* d_own_flags, int32_t n_ranks, uint32_t epoch, void* stream) {
  clear_error();
  if (!d_own_flags || n_ranks < 1 || n_ranks > 8) return fail(MDE_ERR_INVALID, "peer_wait: 1..8 ranks and a flag array");
  PeerFlagsParams p;
  for (int r = 0; r < 8; ++r) p.flags[r] = nullptr;
  p.flags[0] = static_cast<unsigned int*>(d_own_flags);
  p.n_ranks = n_ranks; p.rank = 0; p.epoch = epoch;
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_peer_signal_counter(void* const* d_flags_of_every_rank, int32_t n_ranks, int32_t rank, uint32_t* d_counter, int32_t advance,
                              void* stream) {
  clear_error();
  if (!d_flags_of_every_rank || !d_counter || n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks)
    return fail(MDE_ERR_INVALID, "peer_signal_counter: 1..8 ranks, 0 <= rank < n_ranks, a counter");
  PeerCounterParams p;
  for (int r = 0; r < 8; ++r) p.flags[r] = r < n_ranks ? static_cast<unsigned int*>(d_flags_of_every_rank[r]) : nullptr;
  for (int r = 0; r < n_ranks; ++r)
    if (!p.flags[r]) return fail(MDE_ERR_INVALID, "peer_signal_counter: flag array of rank %d is null", r);
  p.counter = d_counter; p.n_ranks = n_ranks; p.rank = rank; p.advance = advance ? 1 : 0;
  peer_signal_counter_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_peer_wait_counter(void* d_own_flags, int32_t n_ranks, const uint32_t* d_counter, void* stream) {
  clear_error();
  if (!d_own_flags || !d_counter || n_ranks < 1 || n_ranks > 8) return fail(MDE_ERR_INVALID, "peer_wait_counter: 1..8 ranks, a flag array and a counter");
  PeerCounterParams p;
  for (int r = 0; r < 8; ++r) p.flags[r] = nullptr;
  p.flags[0] = static_cast<unsigned int*>(d_own_flags);
  p.counter = const_cast<uint32_t*>(d_counter); p.n_ranks = n_ranks; p.rank = 0; p.advance = 0;
  peer_wait_counter_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_merge_patches(int32_t precision, const void* d_tokens, int32_t per_side, int32_t grid, int32_t padding, int32_t dim,
                        void* d_out, void* stream) {
  clear_error();
  if (!d_tokens || !d_out) return fail(MDE_ERR_INVALID, "merge_patches: null pointer");
  return launch_merge_patches(precision, d_tokens, per_side, grid, padding, dim, d_out, static_cast<cudaStream_t>(stream));
}

int mde_k_qknorm_rope(int32_t precision, void* d_qkv, int64_t rows, int32_t heads, const float* d_qw, const float* d_qb,
                      const float* d_kw, const float* d_kb, float eps, const int32_t* d_pos, const float* d_cos_sin,
                      int32_t max_pos, int32_t gather_n, void* const* d_gather, int32_t gather_ld, void* stream) {
  clear_error();
  if (precision != MDE_FP16 && precision != MDE_BF16) return fail(MDE_ERR_INVALID, "precision must be MDE_FP16 or MDE_BF16");
  if (!d_qkv || !d_qw || !d_qb || !d_kw || !d_kb) return fail(MDE_ERR_INVALID, "qknorm_rope: null pointer");
  return launch_qknorm_rope(precision, d_qkv, rows, heads, d_qw, d_qb, d_kw, d_kb, eps, d_pos, d_cos_sin, max_pos, gather_n, d_gather,
                            gather_ld, static_cast<cudaStream_t>(stream));
}

int mde_k_preprocess_u8_square_pad_cubic(const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w, int32_t dst_h, int32_t dst_w,
                                         int32_t swap_rb, int32_t pad_value, float* d_nchw, void* stream) {
  clear_error();
  if (!d_src || !d_nchw) return fail(MDE_ERR_INVALID, "preprocess_cubic: null pointer");
  if (batch < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return fail(MDE_ERR_INVALID, "preprocess_cubic: empty problem");
  if (dst_h > 65535 || batch * 3 > 65535) return fail(MDE_ERR_INVALID, "preprocess_cubic: output exceeds grid limits");
  if (pad_value < 0 || pad_value > 255) return fail(MDE_ERR_INVALID, "preprocess_cubic: the pad value is an 8-bit level");
  CubicPadParams p;
  p.src = d_src; p.out = d_nchw; p.B = batch; p.src_h = src_h; p.src_w = src_w; p.dst_h = dst_h; p.dst_w = dst_w;
  const int m = std::max(src_h, src_w);
  p.left = (m - src_w) / 2; p.top = (m - src_h) / 2;               // the same count on both sides (core/preprocess.py:239-243)
  p.pad_h = src_h + 2 * p.top; p.pad_w = src_w + 2 * p.left;
  p.swap_rb = swap_rb ? 1 : 0; p.pad_value = pad_value; p.tail = (dst_w * 3) % 8;
  p.scale_y = static_cast<double>(p.pad_h) / dst_h; p.scale_x = static_cast<double>(p.pad_w) / dst_w;
  dim3 grid((dst_w + 255) / 256, dst_h, batch * 3);
  preprocess_cubic_pad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_preprocess_u8_cubic_f32(const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w, int32_t dst_h, int32_t dst_w,
                                  int32_t swap_rb, const double* mean3, const double* std3, float* d_nchw, void* stream) {
  clear_error();
  if (!d_src || !d_nchw) return fail(MDE_ERR_INVALID, "preprocess_cubic_f32: null pointer");
  if (batch < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return fail(MDE_ERR_INVALID, "preprocess_cubic_f32: empty problem");
  if (dst_h > 65535 || batch * 3 > 65535) return fail(MDE_ERR_INVALID, "preprocess_cubic_f32: output exceeds grid limits");
  if ((mean3 == nullptr) != (std3 == nullptr)) return fail(MDE_ERR_INVALID, "preprocess_cubic_f32: mean and std go together (both NULL: the 0..1 image as it is)");
  CubicF32Params p;
  p.src = d_src; p.out = d_nchw; p.B = batch; p.src_h = src_h; p.src_w = src_w; p.dst_h = dst_h; p.dst_w = dst_w;
  p.swap_rb = swap_rb ? 1 : 0; p.tail = (dst_w * 3) % 4;
  p.scale_y = static_cast<double>(src_h) / dst_h; p.scale_x = static_cast<double>(src_w) / dst_w;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3 ? mean3[c] : 0.0; p.std[c] = std3 ? std3[c] : 1.0;
    if (!(p.std[c] > 0.0)) return fail(MDE_ERR_INVALID, "preprocess_cubic_f32: std must be positive");
  }
  dim3 grid((dst_w + 255) / 256, dst_h, batch * 3);
  preprocess_cubic_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MDE_CUDA_TRY(cudaGetLastError());
  return MDE_OK;
}

int mde_k_resize_crops(const void* d_src, int32_t src_is_u8_hwc, int32_t swap_rb, int32_t src_h, int32_t src_w,
                       const mde_crop* crops, int32_t n_crops, int32_t out_h, int32_t out_w, const float* mean3,
                       const float* std3, float* d_out, void* stream) {
  clear_error();
  if (!d_src || !d_out || !crops) return fail(MDE_ERR_INVALID, "resize_crops: null pointer");
  return launch_resize_crops(d_src, src_is_u8_hwc, swap_rb, src_h, src_w, crops, n_crops, out_h, out_w, mean3, std3, d_out,
                             static_cast<cudaStream_t>(stream));
}

int mde_k_depth_pro_post(const float* d_inv, const float* d_fov_deg, int32_t h, int32_t w, int32_t src_h, int32_t src_w,
                         float* d_depth, float* d_f_px, void* stream) {
  clear_error();
  if (!d_inv || !d_fov_deg || !d_depth) return fail(MDE_ERR_INVALID, "depth_pro_post: null pointer");
  return launch_depth_pro_post(d_inv, d_fov_deg, h, w, src_h, src_w, d_depth, d_f_px, static_cast<cudaStream_t>(stream));
}

int mde_k_resize_depth_halfpixel_nan(const float* d_in, int32_t pitch, int32_t h, int32_t w, float* d_out, int32_t out_h, int32_t out_w,
                                     float floor_value, void* stream) {
  clear_error();
  if (!d_in || !d_out) return fail(MDE_ERR_INVALID, "resize_depth_halfpixel_nan: null pointer");
  return launch_resize_depth_halfpixel(d_in, pitch, h, w, d_out, out_h, out_w, 1.f, floor_value, 3.402823466e38f,
                                       static_cast<cudaStream_t>(stream), 1);
}

int mde_k_resize_depth_halfpixel(const float* d_in, int32_t pitch, int32_t h, int32_t w, float* d_out, int32_t out_h, int32_t out_w,
                                 float mul, float clamp_lo, float clamp_hi, void* stream) {
  clear_error();
  if (!d_in || !d_out) return fail(MDE_ERR_INVALID, "resize_depth_halfpixel: null pointer");
  return launch_resize_depth_halfpixel(d_in, pitch, h, w, d_out, out_h, out_w, mul, clamp_lo, clamp_hi, static_cast<cudaStream_t>(stream));
}

int mde_k_resize_depth(const float* d_in, int32_t batch, int32_t hi, int32_t wi, float* d_out, int32_t ho, int32_t wo,
                       float clamp_lo, float clamp_hi, void* stream) {
  clear_error();
  if (!d_in || !d_out) return fail(MDE_ERR_INVALID, "resize_depth: null pointer");
  return launch_resize_depth(d_in, batch, hi, wi, d_out, ho, wo, clamp_lo, clamp_hi, static_cast<cudaStream_t>(stream));
}

int mde_k_upconv_head(int32_t precision, const void* d_z, int32_t ldz, int32_t batch, int32_t hs, int32_t ws, int32_t ho,
                      int32_t wo, const float* d_bias, const float* d_head_w, float head_b, float head_scale,
                      float* d_out, void* stream) {
  clear_error();
  if (!d_z || !d_bias || !d_head_w || !d_out) return fail(MDE_ERR_INVALID, "upconv_head: null pointer");
  return launch_upconv_head(precision, d_z, ldz, batch, hs, ws, ho, wo, d_bias, d_head_w, head_b, head_scale, d_out,
                            static_cast<cudaStream_t>(stream));
}

int mde_k_im2col_f32(int32_t precision, const float* d_nchw, int32_t batch, int32_t h, int32_t w, int32_t patch,
                     int32_t kpad, void* d_cols, void* stream) {
  clear_error();
  return launch_im2col_f32(precision, d_nchw, batch, h, w, patch, kpad, d_cols, static_cast<cudaStream_t>(stream));
}

static int preprocess_entry(int32_t precision, const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w,
                            int32_t dst_h, int32_t dst_w, int32_t patch, int32_t kpad, int32_t swap_rb,
                            const double* mean3, const double* std3, void* d_cols, float* d_nchw, void* stream,
                            int keep_ratio_pad, const double* pad_rgb);

int mde_k_preprocess_u8(int32_t precision, const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w,
                        int32_t dst_h, int32_t dst_w, int32_t patch, int32_t kpad, int32_t swap_rb,
                        const double* mean3, const double* std3, void* d_cols, float* d_nchw, void* stream) {
  clear_error();
  if (!mean3 || !std3) return fail(MDE_ERR_INVALID, "mean/std are required");
  return preprocess_entry(precision, d_src, batch, src_h, src_w, dst_h, dst_w, patch, kpad, swap_rb, mean3, std3, d_cols, d_nchw,
                          stream, 0, nullptr);
}

int mde_k_preprocess_u8_pad(int32_t precision, const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w,
                            int32_t dst_h, int32_t dst_w, int32_t patch, int32_t kpad, int32_t swap_rb,
                            const double* pad_rgb3, const double* mean3, const double* std3, void* d_cols, float* d_nchw,
                            void* stream) {
  clear_error();
  if (!pad_rgb3) return fail(MDE_ERR_INVALID, "the pad colour is required");
  if ((mean3 == nullptr) != (std3 == nullptr)) return fail(MDE_ERR_INVALID, "mean and std go together (both NULL: no normalisation)");
  return preprocess_entry(precision, d_src, batch, src_h, src_w, dst_h, dst_w, patch, kpad, swap_rb, mean3, std3, d_cols, d_nchw,
                          stream, 1, pad_rgb3);
}

static int preprocess_entry(int32_t precision, const uint8_t* d_src, int32_t batch, int32_t src_h, int32_t src_w,
                            int32_t dst_h, int32_t dst_w, int32_t patch, int32_t kpad, int32_t swap_rb,
                            const double* mean3, const double* std3, void* d_cols, float* d_nchw, void* stream,
                            int keep_ratio_pad, const double* pad_rgb) {
  float lut[768];
  if (mean3) build_norm_lut(mean3, std3, lut);
  else build_raw_lut(lut);
  float* d_lut = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MDE_CUDA_TRY(cudaMalloc(&d_lut, sizeof(lut)));
  cudaError_t e = cudaMemcpyAsync(d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, s);
  int rc = MDE_OK;
  if (e != cudaSuccess) rc = fail(MDE_ERR_CUDA, "LUT upload: %s", cudaGetErrorString(e));
  if (rc == MDE_OK)
    rc = launch_preprocess_u8(precision, d_src, static_cast<long long>(src_h) * src_w * 3, batch, src_h, src_w, dst_h,
                              dst_w, patch, kpad, swap_rb, d_lut, d_cols, d_nchw, s, keep_ratio_pad, pad_rgb);
  cudaStreamSynchronize(s);   // test entry point: the temporary LUT must outlive the kernel
  cudaFree(d_lut);
  return rc;
}

}  // extern "C"
