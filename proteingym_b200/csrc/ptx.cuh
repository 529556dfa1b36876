// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM), ldmatrix / mma.sync.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load: coordinates (c0 = innermost element index, c1 = row index); completes on `bar` with the box bytes.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// L2 prefetch of a 2D tile (no shared memory, no barrier): the later cp.async.bulk.tensor load of the same box then hits in L2.
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}

// 2D tiled store smem -> global (bulk async group); out-of-bounds parts of the box are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 2D tiled reduction global[...] += smem[...] (fp32 add performed at L2; one add per element, so deterministic here).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t cvt_f16x2_rn(float lo_elem, float hi_elem) {  // packed RN convert {hi_elem, lo_elem}
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

// two fp32 -> packed e4m3 pair (round to nearest even, saturating at +-448): byte 0 = e0, byte 1 = e1
__device__ __forceinline__ uint32_t cvt_e4m3x2(float e0, float e1) {
  uint16_t d;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(e1), "f"(e0));
  return d;
}
__device__ __forceinline__ uint32_t pack4_e4m3(float e0, float e1, float e2, float e3) {
  return cvt_e4m3x2(e0, e1) | (cvt_e4m3x2(e2, e3) << 16);
}
// ---- packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 process two fp32 lanes per issue slot) and the 3-input max (FMNMX3).
// A pair lives in a 64-bit register {lo = .x, hi = .y}; the pack / unpack moves are register aliasing, not instructions, when the
// two halves are adjacent (tcgen05.ld results are).
__device__ __forceinline__ uint64_t f2_pack(float x, float y) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y));
  return d;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Register re-budgeting between warpgroups (all four warps of a warpgroup execute the same instruction).
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16/bf16 operands, fp32 accumulate). One thread issues.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with kind::f8f6f4 (e4m3 operands as packed bytes in shared memory, K = 32 per instruction, fp32 accumulate; SASS UTCQMMA).
// The instruction descriptor has the layout of make_idesc_f16 with a_format = b_format = 0 meaning E4M3.
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp (the same lane every time for a full warp): the issuer of tcgen05.mma / tcgen05.commit. Keeping the
// surrounding control flow warp-uniform and predicating only the issue lets the compiler hold descriptors in uniform registers;
// a loop that runs under `if (lane == 0)` instead pays an ELECT / branch sequence around every uniform-datapath instruction, and the
// single issuing thread becomes the bottleneck (ncu r02: ~240 SASS instructions per 512-cycle k-block in the GEMM's MMA loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred)::"memory");
  return pred != 0;
}
// Four MMAs over one 128-byte-wide k-block (descriptor start advanced by 32 bytes = +2 in the >>4 field each), in ONE asm block:
// KIND 0 = kind::f16 (4 x K16), KIND 1 = kind::f8f6f4 (4 x K32); CG = cta_group (2: the instruction spans the CTA pair, M = 256,
// issued by the leader CTA only). `accumulate` applies to the first MMA; the other three accumulate.
#define PG_UMMA_KBLOCK(CGS, KINDS)                                                                                                   \
  asm volatile(                                                                                                                      \
      "{\n\t"                                                                                                                        \
      ".reg .pred p, t;\n\t"                                                                                                         \
      ".reg .b64 a1, b1, a2, b2, a3, b3;\n\t"                                                                                        \
      "setp.ne.b32 p, %4, 0;\n\t"                                                                                                    \
      "setp.eq.b32 t, %4, %4;\n\t"                                                                                                   \
      "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\tadd.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\tadd.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t" \
      "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], %1, %2, %3, p;\n\t"                                                       \
      "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], a1, b1, %3, t;\n\t"                                                       \
      "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], a2, b2, %3, t;\n\t"                                                       \
      "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], a3, b3, %3, t;\n\t"                                                       \
      "}" ::"r"(tmem_d),                                                                                                             \
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)                                                                            \
      : "memory")
template <int KIND, int CG = 1>
__device__ __forceinline__ void umma_kblock(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    if (KIND == 0) PG_UMMA_KBLOCK("1", "f16"); else PG_UMMA_KBLOCK("1", "f8f6f4");
  } else {
    if (KIND == 0) PG_UMMA_KBLOCK("2", "f16"); else PG_UMMA_KBLOCK("2", "f8f6f4");
  }
}
#undef PG_UMMA_KBLOCK

// ---------------------------------------------------------------- CTA pairs (cta_group::2): forms taken from CUTLASS' sm100 headers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared address: names the LEADER CTA's copy of a variable
// 2D tiled load issued by either CTA of a pair into ITS OWN shared memory; the transaction bytes complete on the leader CTA's barrier.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// All previously issued cta_group::2 MMAs of this thread arrive on `bar` in BOTH CTAs of the pair when complete.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster. Default semantics (.release.cta), as for a
// local arrive: the only thing this arrive publishes is that the caller's tcgen05.ld of a TMEM buffer has completed (wait::ld +
// tcgen05.fence::before_thread_sync precede it), no generic-proxy memory. The earlier `.release.cluster` form compiled to
// MEMBAR.ALL.GPU + ERRBAR in front of every remote arrive — 3 % of all warp samples of the fc1 GEMM sat in that fence
// (profiles/prof_r02_gemm_fc1_f16f8_cta2.ncu-rep), each one holding back the release of an accumulator buffer to the leader's MMA
// thread until the epilogue's outstanding global stores had drained.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {  // one whole warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// A operand from TMEM (fp16 packed), B from smem.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// O (+)= P * V over 64 keys (4 k-steps of 16) in ONE asm block. A = P as packed fp16 pairs in TMEM (8 columns per k-step; hi pairs at
// tmem_p, lo pairs LO columns further), B = V (hi plane at vh, lo plane at vl; descriptor start +128 = 16 key rows of 128 B per
// k-step). NP == 2 issues the three products hi*hi + lo*hi + hi*lo per k-step. `accumulate` applies to the very first MMA.
template <int NP, int LO>
__device__ __forceinline__ void umma_pv64(uint32_t tmem_o, uint32_t tmem_p, uint64_t vh, uint64_t vl, uint32_t idesc, uint32_t accumulate) {
  if (NP == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 a1, a2, a3;\n\t"
        ".reg .b64 b1, b2, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, %4, %4;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.s64 b1, %2, 128;\n\tadd.s64 b2, %2, 256;\n\tadd.s64 b3, %2, 384;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n\t"
        "}" ::"r"(tmem_o),
        "r"(tmem_p), "l"(vh), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 a1, a2, a3, l0, l1, l2, l3;\n\t"
        ".reg .b64 b1, b2, b3, c1, c2, c3;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, %5, %5;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "add.u32 l0, %1, %6;\n\tadd.u32 l1, a1, %6;\n\tadd.u32 l2, a2, %6;\n\tadd.u32 l3, a3, %6;\n\t"
        "add.s64 b1, %2, 128;\n\tadd.s64 b2, %2, 256;\n\tadd.s64 b3, %2, 384;\n\t"
        "add.s64 c1, %3, 128;\n\tadd.s64 c2, %3, 256;\n\tadd.s64 c3, %3, 384;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %4, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [l0], %2, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %3, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [l1], b1, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], c1, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [l2], b2, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], c2, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [l3], b3, %4, t;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], c3, %4, t;\n\t"
        "}" ::"r"(tmem_o),
        "r"(tmem_p), "l"(vh), "l"(vl), "r"(idesc), "r"(accumulate), "n"(LO)
        : "memory");
  }
}
// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100 "version 1"):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
// K-major operand tile stored as rows of 128 bytes (64 fp16) under the 128B swizzle: SBO = 8 rows * 128 B = 1024 B,
// LBO unused for swizzled K-major (set to 1 like CUTLASS does).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: fp16 A/B, fp32 D.
//   [4,6) c_format=1(F32) | [7,10) a_format (0 = F16) | [10,13) b_format | [15] a_major (0=K) | [16] b_major |
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread i gets lane base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: mirror of tmem_ld_32x32b_x32.
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- legacy warp MMA (attention v1)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// fp32 -> fp16 hi (+ residual lo), saturating so an outlier cannot become inf
__device__ __forceinline__ __half f2h_sat(float v) {
  unsigned short u;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(u) : "f"(v));
  return __ushort_as_half(u);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}
__device__ __forceinline__ void split_hi_lo(float v, __half& hi, __half& lo) {
  hi = f2h_sat(v);
  lo = __float2half_rn(v - __half2float(hi));
}

}  // namespace pg
