// sm_100a tensor-core plumbing shared by the convolution kernels: mbarrier, TMA (tiled tensor
// maps), tcgen05 (alloc / mma / commit / ld) wrappers and descriptor builders.
// Encodings follow the PTX ISA tcgen05 matrix/instruction descriptor tables.
#pragma once
#include <cuda.h>
#include <string.h>
#include "common.cuh"

namespace tc {

// ---------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// Warp-uniform role code.  The single-thread instructions of this file (UTMALDG, UTCHMMA, UTCBAR in SASS) take their
// operands from UNIFORM registers.  Issued from an `if (lane == 0)` region the compiler treats every value as per-thread
// and wraps each instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~20 dependent instructions per MMA): the
// issuing thread then needs about as long for its instruction stream as the tensor pipe needs for the MMAs.  The role
// warps therefore run CONVERGED (all 32 lanes walk the pipeline loops, branch conditions made uniform with `uniform()`),
// and only the instruction itself is predicated on the lane `elect_one()` picked: descriptors, coordinates and barrier
// addresses then live in uniform registers and the MMAs of a stage issue back to back.
__device__ __forceinline__ uint32_t elect_one() {          // 1 in exactly one (always the same) lane of a converged warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }   // provably warp-uniform

// ---------------------------------------------------------------------------- cheap tile decoding
// Every pipeline role decodes its next tile from a linear index; with plain `/` and `%` by launch parameters and a loop over the
// taps this took ~600-800 SASS instructions per tile (three to five emulated 32-bit divisions with I2F / MUFU.RCP chains), i.e.
// ~2.5 k cycles in which the MMA issuer issues nothing (clock64 trace: one ~2650-cycle gap per tile, profiles/r02_v16_*).
// Division by a launch constant: q = umulhi(n, mul) >> shr with mul = ceil(2^(31 + ceil(log2 d)) / d), exact for 0 <= n < 2^31
// (round-up multiplier method); d == 1 is encoded as mul == 0.
struct FastDiv { uint32_t mul, shr; };
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f; f.mul = 0; f.shr = 0;
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                       // ceil(log2 d)
  const uint32_t pw = 31 + l;
  f.mul = (uint32_t)((((uint64_t)1 << pw) + d - 1) / d);
  f.shr = pw - 32;
  return f;
}
__host__ __device__ __forceinline__ uint32_t fast_div(uint32_t n, FastDiv f) {
  if (f.mul == 0) return n;
#ifdef __CUDA_ARCH__
  return __umulhi(n, f.mul) >> f.shr;
#else
  return (uint32_t)(((uint64_t)n * f.mul) >> 32) >> f.shr;
#endif
}
// Which taps of a tile touch the image at all (the others lie in the padding and are skipped) depends on the tile's row and on
// its column separately: bit i of rows[ht] & cols[wt] = tap i contributes.  Built on the host when both tile counts fit.
constexpr int TAP_TABLE = 256;
struct TapTables { uint16_t rows[TAP_TABLE], cols[TAP_TABLE]; int on; };

// ---------------------------------------------------------------------------- programmatic dependent launch
// A kernel launched with the programmatic-stream-serialization attribute (tc::launch below) may start while its predecessor
// in the stream is still running: its CTAs become resident as the predecessor's CTAs retire, run their prologue (barrier
// init, TMEM allocation, tensor-map prefetch -- nothing that touches global memory) and block in pdl_wait() until the
// predecessor grid has completed and its memory operations are visible.  pdl_launch_dependents() is the predecessor's
// side: "my dependents may be scheduled now" (they still wait for this grid's completion in their pdl_wait()).
// Both are no-ops in a launch without the attribute / without dependents.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------- TMA
// bulk L2 prefetch of `bytes` (multiple of 16) contiguous bytes at a 16 B aligned global address
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// L2 prefetch of a 4-D box (no shared-memory destination, no completion mechanism: a hint)
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = lane, reg j = column j).
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Instruction descriptor, kind::tf32, fp32 accumulate (PTX ISA "Instruction descriptor" table):
//  [4,6) D format (1 = F32)  [7,10) A format (2 = TF32)  [10,13) B format (2 = TF32)
//  [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t m, uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// Shared-memory matrix descriptor (PTX ISA "Matrix descriptor"): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version 1 at bit 46, swizzle mode [61,64) (2 = 128B).
// layout_type: 2 = SWIZZLE_128B (16 B atoms; K-major operands), 1 = SWIZZLE_128B_BASE32B (32 B atoms: the only
// swizzled layout the tensor core accepts for MN-major tf32 operands; TMA side = SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled();

// fp32 tensor, 128B swizzle, zero OOB fill.  dims/strides innermost first; strides_bytes has rank-1 entries.
int make_tmap_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides, bool atom32 = false);
// the same without swizzle: boxes with more than 128 B per row, used for L2 prefetches only (nothing lands in shared memory)
int make_tmap_f32_linear(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, const uint32_t* elem_strides);

// 1 = launch the tensor-core kernels with the programmatic-stream-serialization attribute (environment B200SEG_PDL or
// b2_debug_set(11, v)); 0 = plain stream order
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}

}  // namespace tc
