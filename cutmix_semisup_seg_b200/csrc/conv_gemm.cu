// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, kind::tf32, fp32 accumulate in
// TMEM) fed by TMA.  Serves fprop and dgrad of every nn.Conv2d on the hot path
// (architectures/deeplab2.py:65-128,140-150; torchvision ResNet-101 / ASPP and
// architectures/deeplab3plus.py:29-48 for DeepLab v3+).
//
// GEMM view:  D[pixel, n] = sum_{tap} sum_{k} A[pixel @ tap, k] * B[n, tap, k]
//   * M tile  = 128 output pixels = a (BW x BH x BN) box of the NHWC activation tensor.  The TMA
//     engine gathers the box for a filter tap by shifting the start coordinate by (dh, dw); pixels
//     outside the image are zero-filled by TMA (= conv padding), strided convs use TMA element strides.
//     No im2col buffer ever exists in HBM.
//   * N tile  <= 256 output channels, K block = 32 channels (one 128 B swizzle row).
//   * 4-stage smem ring (16 KB A + 32 KB B per stage), 2 x 256-column TMEM accumulators, so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
//   * persistent CTAs (one per SM), warp-specialised: warp0 TMA producer, warp1 MMA issuer
//     (single elected thread), warp2 TMEM allocator, warps 4-7 epilogue (TMEM -> regs -> HBM).
//   * taps whose box lies completely in the zero padding are skipped (ASPP dilation 12/24/36).
//   * fused epilogue: BN scale/shift | bias, residual add, ReLU, ReLU-gate (backward), second scale,
//     accumulate; output leading dimension + channel offset let a conv write into a slice of a wider
//     buffer (no torch.cat), output stride/offset scatter handles strided dgrad.
#include <stdlib.h>
#include "tc_common.cuh"
#include "conv_epilogue.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                       // fp32 elements = 128 bytes
constexpr int MAX_BLOCK_N = 224;                  // wider N tiles run on CTA pairs (conv_gemm2.cu)
constexpr int ACC_COLS = 256;                     // TMEM columns per accumulator stage
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 4;      // 16 KB
constexpr int B_STAGE_BYTES = MAX_BLOCK_N * BLOCK_K * 4;  // 32 KB
constexpr int EPI_BYTES = epi::BYTES;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
constexpr int MAX_TAPS = 16;
constexpr int NUM_THREADS = 384;      // 4 control warps + 8 epilogue warps
constexpr int TMEM_COLS = 512;

struct ConvKArgs {
  int n, ih, iw, k;
  int nb;
  int oh, ow;
  int fh, fw, ldd, ostride, ooh, oow;
  int istride;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles_n, block_n;
  int num_tiles;
  int n_taps;
  short dh[MAX_TAPS], dw[MAX_TAPS], btap[MAX_TAPS];
  int kblocks;
  int n_pass;
  int tap_outer;             // 0: K-block outer / tap inner (default, see the producer); 1: tap outer (debug knob 7)
  int reverse;               // 1: walk the tiles from the last to the first (alternating launch directions, see b2_conv_gemm)
  tc::FastDiv fd_w, fd_h, fd_nn;   // division by tiles_w, tiles_h, n_tiles_n (tc_common.cuh "cheap tile decoding")
  tc::TapTables tt;
  epi::Params ep;             // epilogue parameter block (conv_epilogue.cuh)
};

struct TileInfo {
  int n_idx, w0, h0, n0;
  int m_idx;                  // index of the M tile (row of the fused column statistics)
  uint32_t tap_mask;
};

__device__ __forceinline__ TileInfo decode_tile(const ConvKArgs& a, int tile) {
  TileInfo t;
  if (a.reverse) tile = a.num_tiles - 1 - tile;
  const int m = (int)tc::fast_div((uint32_t)tile, a.fd_nn);
  t.n_idx = tile - m * a.n_tiles_n;
  t.m_idx = m;
  const int q = (int)tc::fast_div((uint32_t)m, a.fd_w);
  const int wt = m - q * a.tiles_w;
  const int nt = (int)tc::fast_div((uint32_t)q, a.fd_h);
  const int ht = q - nt * a.tiles_h;
  t.w0 = wt * a.bw; t.h0 = ht * a.bh; t.n0 = nt * a.bn;
  uint32_t mask = 0;
  if (a.tt.on) {
    mask = (uint32_t)a.tt.rows[ht] & (uint32_t)a.tt.cols[wt];
  } else {
    for (int i = 0; i < a.n_taps; ++i) {
      const int lo_h = t.h0 * a.istride + a.dh[i], hi_h = lo_h + (a.bh - 1) * a.istride;
      const int lo_w = t.w0 * a.istride + a.dw[i], hi_w = lo_w + (a.bw - 1) * a.istride;
      if (hi_h >= 0 && lo_h < a.ih && hi_w >= 0 && lo_w < a.iw) mask |= 1u << i;
    }
  }
  if (mask == 0) mask = 1;  // accumulator must still be written (all-zero contribution)
  t.tap_mask = mask;
  return t;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                 const __grid_constant__ ConvKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = (int)tc::uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;     // uniform: role branches converge

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmB);
    if (a.n_pass > 1) { tc::tma_prefetch_desc(&tmAlo); tc::tma_prefetch_desc(&tmBlo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull_bar[i], 1); tc::mbar_init(&tempty_bar[i], 8); }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_ptr, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  tc::pdl_wait();                 // everything above overlaps the predecessor's tail (tc_common.cuh)
  tc::pdl_launch_dependents();

  const uint32_t rows_a = a.bw * a.bh * a.bn;
  const uint32_t stage_tx = rows_a * 128u + (uint32_t)a.block_n * 128u;

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues; tc_common.cuh) =====================
    const uint32_t el = tc::elect_one();
    int stage = 0; uint32_t phase = 0;
    // K-block outer / tap inner: the shifted boxes of one 32-channel slab are served from L2 (see conv_gemm2.cu)
    const int n_outer = a.tap_outer ? a.n_taps : a.kblocks;
    const int n_inner = a.tap_outer ? a.kblocks : a.n_taps;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const TileInfo t = decode_tile(a, tile);
      const int w_base = t.w0 * a.istride, h_base = t.h0 * a.istride;
      for (int o = 0; o < n_outer; ++o) {
        for (int i = 0; i < n_inner; ++i) {
          const int tap = a.tap_outer ? o : i;
          const int kb = a.tap_outer ? i : o;
          if (!(t.tap_mask >> tap & 1)) continue;
          const int cw = w_base + a.dw[tap];
          const int ch = h_base + a.dh[tap];
          const int bt = a.btap[tap];
          for (int p = 0; p < a.n_pass; ++p) {
            tc::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (el) {
              tc::mbar_expect_tx(&full_bar[stage], stage_tx);
              tc::tma_load_4d(smem_a + stage * A_STAGE_BYTES, (p & 1) ? &tmAlo : &tmA, &full_bar[stage],
                              kb * BLOCK_K, cw, ch, t.n0);
              tc::tma_load_3d(smem_b + stage * B_STAGE_BYTES, (p & 2) ? &tmBlo : &tmB, &full_bar[stage],
                              kb * BLOCK_K, bt, t.n_idx * a.block_n);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    const uint32_t el = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_tf32(BLOCK_M, a.block_n, 0, 0);
    // K-major SWIZZLE_128B descriptors: low word = start address >> 4 | LBO << 16, high word constant; one K step of
    // 8 tf32 = 32 B = +2 in the low word
    const uint32_t desc_hi = (uint32_t)(tc::make_smem_desc_sw128(0, 16, 1024) >> 32);
    const uint32_t a_lo0 = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(smem_a), 16, 1024);
    const uint32_t b_lo0 = (uint32_t)tc::make_smem_desc_sw128(tc::smem_u32(smem_b), 16, 1024);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const TileInfo t = decode_tile(a, tile);
      tc::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
      uint32_t accum = 0;
      const int iters = __popc(t.tap_mask) * a.kblocks * a.n_pass;
      for (int it = 0; it < iters; ++it) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (A_STAGE_BYTES >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)stage * (B_STAGE_BYTES >> 4);
        if (el) {
#pragma unroll
          for (int ks = 0; ks < BLOCK_K / 8; ++ks)
            tc::mma_tf32(tmem_d, ((uint64_t)desc_hi << 32) | (a_lo + 2 * ks), ((uint64_t)desc_hi << 32) | (b_lo + 2 * ks), idesc,
                         ks == 0 ? accum : 1u);
          tc::mma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
        }
        accum = 1;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (el) tc::mma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (conv_epilogue.cuh) =====================
    const int ew = (warp - 4) & 3;        // TMEM lane quarter
    const int eh = (warp - 4) >> 2;       // even / odd chunks
    const int ewi = warp - 4;
    const int row = ew * 32 + lane;
    float* stg = reinterpret_cast<float*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256) + ewi * (32 * epi::ROW_FLOATS);
    int* rowpix = reinterpret_cast<int*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + epi::NUM_WARPS * epi::WARP_BYTES) + ewi * 32;
    const epi::Params& ep = a.ep;     // stays in the kernel's constant parameter space
    int acc = 0; uint32_t acc_phase = 0;
    const int bwbh = a.bw * a.bh;
    const int dn = row / bwbh;
    const int rem = row - dn * bwbh;
    const int dhh = rem / a.bw;
    const int dww = rem - dhh * a.bw;
    auto row_pixel = [&](const TileInfo& ti) -> int {       // output pixel index of this thread's accumulator row
      const int pn = ti.n0 + dn, ph = ti.h0 + dhh, pw = ti.w0 + dww;
      const bool valid = (uint32_t)row < rows_a && pn < a.n && ph < a.oh && pw < a.ow;
      const long long pix = ((long long)pn * a.fh + (long long)ph * a.ostride + a.ooh) * a.fw + (long long)pw * a.ostride + a.oow;
      return valid ? (int)pix : -1;
    };
    const bool has_reads = (ep.dbg & 4) && (ep.addend || ep.gate || ep.accumulate);
    const int half_cols = ((a.block_n / 32 + 1) / 2) * 32;      // the two warps of a lane quarter prefetch half a row each
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const TileInfo t = decode_tile(a, tile);
      __syncwarp();
      rowpix[lane] = row_pixel(t);
      __syncwarp();
      if (has_reads && tile + (int)gridDim.x < a.num_tiles) {   // next tile's epilogue operands -> L2 (see prefetch_row)
        const TileInfo tn = decode_tile(a, tile + gridDim.x);
        epi::prefetch_row(ep, row_pixel(tn), tn.n_idx * a.block_n + eh * half_cols, eh ? a.block_n - half_cols : half_cols);
      }
      tc::mbar_wait(&tfull_bar[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * ACC_COLS;
      const int stat_row = t.m_idx * 4 + ew;                   // fused column statistics: (M tile, lane quarter)
      epi::drain_tile(ep, taddr, a.block_n, t.n_idx * a.block_n, stg, rowpix, lane, eh, stat_row, 0u, [&]() {
        tc::tc_fence_before();           // accumulator fully read: hand the TMEM stage back to the MMA warp
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);
      });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// ------------------------------------------------------------------------------------------ host
namespace tc {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides, bool atom32) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return b2_fail(B2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return b2_fail(B2_ERR_INVALID, "tensor map base %p not 16B aligned", ptr);
  for (int i = 0; i < rank - 1; ++i)
    if (strides_bytes[i] & 15) return b2_fail(B2_ERR_INVALID, "tensor map stride[%d]=%llu not a multiple of 16 B", i, (unsigned long long)strides_bytes[i]);
  for (int i = 0; i < rank; ++i)
    if (box[i] == 0 || box[i] > 256) return b2_fail(B2_ERR_INVALID, "tensor map box[%d]=%u out of range", i, box[i]);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                   elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return b2_fail(B2_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return B2_OK;
}

int make_tmap_f32_linear(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return b2_fail(B2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                   elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return b2_fail(B2_ERR_CUDA, "cuTensorMapEncodeTiled (linear) failed with CUresult %d", (int)r);
  return B2_OK;
}

}  // namespace tc

// Pick the (BW, BH, BN) pixel box of at most `max_rows` rows that wastes the fewest tile slots.
void b2_choose_box(int ow, int oh, int n, int max_rows, int istride, int* bw_o, int* bh_o, int* bn_o) {
  double best = -1.0; int bbw = 1, bbh = 1, bbn = 1;
  const int max_b = 256 / istride;
  for (int bw = 1; bw <= ow && bw <= max_rows && bw <= max_b; ++bw) {
    for (int bh = 1; bh <= oh && bw * bh <= max_rows && bh <= max_b; ++bh) {
      int bn = 1;
      if (bw == ow && bh == oh) { bn = max_rows / (bw * bh); if (bn > n) bn = n; if (bn < 1) bn = 1; }
      const double tiles = (double)((ow + bw - 1) / bw) * ((oh + bh - 1) / bh) * ((n + bn - 1) / bn);
      const double eff = ((double)ow * oh * n) / (tiles * max_rows);
      // prefer wide boxes on ties (longer contiguous runs per TMA row group)
      const double score = eff + 1e-6 * bw;
      if (score > best) { best = score; bbw = bw; bbh = bh; bbn = bn; }
    }
  }
  *bw_o = bbw; *bh_o = bbh; *bn_o = bbn;
}

int b2_conv_gemm_2cta(const b2_conv_params* p, void* stream);   // conv_gemm2.cu

// Row blocks of the fused-statistics buffer: 4 (TMEM lane quarters) per 128-pixel M tile; same tiling in both kernels.
extern "C" int64_t b2_conv_stats_rows(const b2_conv_params* p) {
  if (!p || p->ow < 1 || p->oh < 1 || p->n < 1 || p->istride < 1) return 0;
  int bw, bh, bn;
  b2_choose_box(p->ow, p->oh, p->n, 128, p->istride, &bw, &bh, &bn);
  const int64_t m_tiles = (int64_t)((p->ow + bw - 1) / bw) * ((p->oh + bh - 1) / bh) * ((p->n + bn - 1) / bn);
  return m_tiles * 4;
}
int g_conv_force_1cta = 0;
int g_conv_epi_debug = 0;
int g_conv_pdl = -1;              // b2_debug_set(11, v) / environment B200SEG_PDL: programmatic dependent launch (tc_common.cuh)
bool tc::pdl_enabled() {
  if (g_conv_pdl < 0) { const char* e = getenv("B200SEG_PDL"); g_conv_pdl = e ? atoi(e) : 0; }
  return g_conv_pdl > 0;
}
// Alternating tile directions (b2_debug_set(12, v) / environment B200SEG_ALT_DIR): consecutive fprop / dgrad launches walk their
// tiles in opposite directions.  A layer's input is usually the tensor the previous launch has just written; the ~100 MB it wrote
// last are still in the 126 MB L2, so the consumer starts where the producer stopped instead of at the other end of the tensor
// (whose L2 lines the producer itself has long evicted).  Results do not depend on the tile order.
int g_conv_alt_dir = -1;
int g_conv_next_reverse = 0;      // direction of the launch being prepared (set by b2_conv_gemm)
static int g_conv_dir_state = 0;
// b2_debug_set(18, v) / environment B200SEG_WIDE_PF: the CTA-pair kernel prefetches the A stream of 1x1 layers with K >= 512 into
// L2 in boxes of 256 channels (1 KB per pixel row instead of the 128 B rows of the operand boxes): DRAM-page-friendly reads
int g_conv_wide_pf = -1;
int g_conv_tap_tables = 1;        // b2_debug_set(17, 0): per-tile tap masks by the loop over the taps instead of the host-built tables (A/B)
int g_conv_tap_outer = 0;         // b2_debug_set(7, 1): producer loops tap-outer / K-block-inner (the round-1 order)

extern "C" int b2_conv_gemm(const b2_conv_params* p, void* stream) {
  B2_REQUIRE(p && p->a && p->b && p->d, "b2_conv_gemm: null tensor");
  B2_REQUIRE(p->n > 0 && p->ih > 0 && p->iw > 0 && p->k > 0 && p->nb > 0 && p->oh > 0 && p->ow > 0, "b2_conv_gemm: bad dims");
  B2_REQUIRE(p->n_taps >= 1 && p->n_taps <= MAX_TAPS && p->taps, "b2_conv_gemm: n_taps=%d out of range", p->n_taps);
  B2_REQUIRE(p->lda % 4 == 0 && p->lda >= p->k, "b2_conv_gemm: lda=%d must be a multiple of 4 and >= K", p->lda);
  B2_REQUIRE(p->ldb % 4 == 0 && p->ldb >= p->k, "b2_conv_gemm: ldb=%d must be a multiple of 4 and >= K", p->ldb);
  B2_REQUIRE(p->istride >= 1 && p->istride <= 8 && p->ostride >= 1, "b2_conv_gemm: bad strides");
  B2_REQUIRE(p->n_split == 1 || p->n_split == 3 || p->n_split == 4, "b2_conv_gemm: n_split must be 1, 3 or 4");
  B2_REQUIRE(p->n_split == 1 || (p->a_lo && p->b_lo), "b2_conv_gemm: split mode needs a_lo and b_lo");
  B2_REQUIRE(p->ldd >= 1 && p->fh >= 1 && p->fw >= 1, "b2_conv_gemm: bad output geometry");

  for (int i = 0; i < p->n_taps; ++i)
    B2_REQUIRE(p->taps[i * 3 + 2] >= 0 && p->taps[i * 3 + 2] < p->tb, "b2_conv_gemm: tap %d references weight tap %d >= %d", i, p->taps[i * 3 + 2], p->tb);
  if (p->stats) {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    B2_REQUIRE(p->gate && !p->accumulate && p->nb % 4 == 0 && p->ldd % 4 == 0 && p->ld_gate % 4 == 0 && al16(p->d) && al16(p->gate) &&
               (!p->addend || (p->ld_add % 4 == 0 && al16(p->addend))) && p->ld_stats % 4 == 0 && p->ld_stats >= p->nb && al16(p->stats) &&
               (!p->stats_sub || (p->ld_stats_sub % 4 == 0 && al16(p->stats_sub))),
               "b2_conv_gemm: fused statistics need a gate, no accumulate, and 16 B aligned / 4-float-padded operands");
  }
  if (g_conv_alt_dir < 0) { const char* e = getenv("B200SEG_ALT_DIR"); g_conv_alt_dir = e ? atoi(e) : 0; }
  g_conv_next_reverse = 0;
  if (g_conv_alt_dir > 0) { g_conv_next_reverse = g_conv_dir_state; g_conv_dir_state ^= 1; }
  // N tiles of 256 output channels run on CTA pairs (tcgen05 cta_group::2): see conv_gemm2.cu
  if (p->nb > 224 && !g_conv_force_1cta && b2_sm_count_cached() >= 2 && (p->max_ctas == 0 || p->max_ctas >= 2))
    return b2_conv_gemm_2cta(p, stream);

  ConvKArgs a;
  memset(&a, 0, sizeof(a));
  a.n = p->n; a.ih = p->ih; a.iw = p->iw; a.k = p->k; a.nb = p->nb; a.oh = p->oh; a.ow = p->ow;
  a.fh = p->fh; a.fw = p->fw; a.ldd = p->ldd; a.ostride = p->ostride; a.ooh = p->ooh; a.oow = p->oow;
  a.istride = p->istride;
  b2_choose_box(p->ow, p->oh, p->n, BLOCK_M, p->istride, &a.bw, &a.bh, &a.bn);
  a.tiles_w = (p->ow + a.bw - 1) / a.bw; a.tiles_h = (p->oh + a.bh - 1) / a.bh; a.tiles_n = (p->n + a.bn - 1) / a.bn;
  int block_n = ((p->nb + 31) / 32) * 32;
  if (block_n > MAX_BLOCK_N) block_n = MAX_BLOCK_N;
  a.block_n = block_n;
  a.n_tiles_n = (p->nb + block_n - 1) / block_n;
  const int64_t num_tiles = (int64_t)a.tiles_w * a.tiles_h * a.tiles_n * a.n_tiles_n;
  B2_REQUIRE(num_tiles < (1ll << 31), "b2_conv_gemm: too many tiles");
  a.num_tiles = (int)num_tiles;
  a.n_taps = p->n_taps;
  for (int i = 0; i < p->n_taps; ++i) {
    a.dh[i] = (short)p->taps[i * 3 + 0]; a.dw[i] = (short)p->taps[i * 3 + 1]; a.btap[i] = (short)p->taps[i * 3 + 2];
    B2_REQUIRE(p->taps[i * 3 + 2] >= 0 && p->taps[i * 3 + 2] < p->tb, "b2_conv_gemm: tap %d references weight tap %d >= %d", i, p->taps[i * 3 + 2], p->tb);
  }
  a.kblocks = (p->k + BLOCK_K - 1) / BLOCK_K;
  a.n_pass = p->n_split;
  a.tap_outer = g_conv_tap_outer;
  a.reverse = g_conv_next_reverse;
  a.fd_w = tc::make_fastdiv((uint32_t)a.tiles_w); a.fd_h = tc::make_fastdiv((uint32_t)a.tiles_h);
  a.fd_nn = tc::make_fastdiv((uint32_t)a.n_tiles_n);
  if (a.tiles_w <= tc::TAP_TABLE && a.tiles_h <= tc::TAP_TABLE && g_conv_tap_tables != 0) {
    for (int ht = 0; ht < a.tiles_h; ++ht) {
      uint32_t mk = 0;
      for (int i = 0; i < a.n_taps; ++i) {
        const int lo = ht * a.bh * a.istride + a.dh[i], hi = lo + (a.bh - 1) * a.istride;
        if (hi >= 0 && lo < a.ih) mk |= 1u << i;
      }
      a.tt.rows[ht] = (uint16_t)mk;
    }
    for (int wt = 0; wt < a.tiles_w; ++wt) {
      uint32_t mk = 0;
      for (int i = 0; i < a.n_taps; ++i) {
        const int lo = wt * a.bw * a.istride + a.dw[i], hi = lo + (a.bw - 1) * a.istride;
        if (hi >= 0 && lo < a.iw) mk |= 1u << i;
      }
      a.tt.cols[wt] = (uint16_t)mk;
    }
    a.tt.on = 1;
  }
  a.ep.d = p->d; a.ep.ldd = p->ldd; a.ep.nb = p->nb;
  a.ep.scale = p->scale; a.ep.shift = p->shift; a.ep.addend = p->addend; a.ep.gate = p->gate; a.ep.scale2 = p->scale2;
  a.ep.ld_add = p->ld_add; a.ep.ld_gate = p->ld_gate; a.ep.relu = p->relu; a.ep.accumulate = p->accumulate;
  a.ep.stats = p->stats; a.ep.ld_stats = p->ld_stats; a.ep.sub = p->stats_sub; a.ep.ld_sub = p->ld_stats_sub;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  a.ep.dbg = g_conv_epi_debug;
  a.ep.vec_ok = (p->ldd % 4 == 0) && al16(p->d) && (!p->addend || (p->ld_add % 4 == 0 && al16(p->addend))) &&
             (!p->gate || (p->ld_gate % 4 == 0 && al16(p->gate)));

  // A: (K, IW, IH, N) fp32, box (32, BW*s, BH*s, BN) with element strides (1, s, s, 1)
  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  {
    const uint64_t dims[4] = {(uint64_t)p->k, (uint64_t)p->iw, (uint64_t)p->ih, (uint64_t)p->n};
    const uint64_t strides[3] = {(uint64_t)p->lda * 4, (uint64_t)p->iw * p->lda * 4, (uint64_t)p->ih * p->iw * p->lda * 4};
    const uint32_t box[4] = {BLOCK_K, (uint32_t)(a.bw * p->istride), (uint32_t)(a.bh * p->istride), (uint32_t)a.bn};
    const uint32_t es[4] = {1, (uint32_t)p->istride, (uint32_t)p->istride, 1};
    int rc = tc::make_tmap_f32(&tmA, p->a, 4, dims, strides, box, es);
    if (rc) return rc;
    rc = tc::make_tmap_f32(&tmAlo, p->a_lo ? p->a_lo : p->a, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)p->k, (uint64_t)p->tb, (uint64_t)p->nb};
    const uint64_t strides[2] = {(uint64_t)p->ldb * 4, (uint64_t)p->tb * p->ldb * 4};
    const uint32_t box[3] = {BLOCK_K, 1, (uint32_t)block_n};
    const uint32_t es[3] = {1, 1, 1};
    int rc = tc::make_tmap_f32(&tmB, p->b, 3, dims, strides, box, es);
    if (rc) return rc;
    rc = tc::make_tmap_f32(&tmBlo, p->b_lo ? p->b_lo : p->b, 3, dims, strides, box, es);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  int grid = b2_sm_count_cached();
  if (grid <= 0) return b2_fail(B2_ERR_CUDA, "b2_conv_gemm: no CUDA device");
  if (p->max_ctas > 0 && p->max_ctas < grid) grid = p->max_ctas;
  if (grid > a.num_tiles) grid = a.num_tiles;
  tc::launch(conv_gemm_kernel, (unsigned)grid, NUM_THREADS, SMEM_BYTES, (cudaStream_t)stream, tmA, tmAlo, tmB, tmBlo, a);
  B2_LAUNCH_CHECK("conv_gemm_kernel");
  return B2_OK;
}
