// K3 — tcgen05/TMA GEMM with fused epilogues for the linear layers of the ESM / Tranception forward
// (reference ops: esm/multihead_attention.py:243-261,395; esm/modules.py:138-140).
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T + bias )        A, W K-major (row-major), fp32 accumulate in TMEM.
//
// Precision (the reference is fp32). Every operand x is an fp16 pair hi = rn(x), lo = rn(x - hi). Three operand modes:
//   nseg 1  single pass            hi*hi                                                     (fast mode, ~1e-2 on scores)
//   nseg 3  fp16 x3                hi*hi + lo*hi + hi*lo, all kind::f16                      (3 tensor-pipe units)
//   nseg 2  fp16 + fp8 cross terms hi*hi (kind::f16) + [lo8 | hi8] * [hi8 | lo8]^T (kind::f8f6f4, e4m3, K-concatenated, 2x rate)
//           with lo8 = e4m3(lo * 2^11 * s), hi8 = e4m3(hi * s): s = a fixed power of two on the A side (applied by the producing
//           kernel), a per-row power of two t_n on the W side (chosen at load time); the epilogue rescales by 1/(2^11 s t_n).
//           2 tensor-pipe units; emulated on CPU at true ESM-1v size: 3.0e-4 max score error (scripts/precision_f8.py).
//
// Accumulation: tcgen05.mma adds into its fp32 TMEM accumulator with truncation, a bias that grows with the number of
// accumulation steps (DESIGN.md "hardware finding"). The K loop is therefore cut into CHUNKS: each chunk accumulates into one of
// the two TMEM buffers from zero, and the epilogue warps add the chunks in registers in round-to-nearest fp32 (fixed order, so
// results stay deterministic). The small-magnitude cross terms form chunk 0 (their own truncation is irrelevant), the hi*hi
// product is split into pieces of <= kchunk (default 1280) along K. The same double-buffered TMEM that used to overlap the
// epilogue of tile i with the MMAs of tile i+1 now also overlaps the read-out of chunk c with the MMAs of chunk c+1.
//
// Structure (one persistent CTA per SM, 384 threads, no clusters):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D tiles (128 rows x 128 B of A, 256 rows x 128 B of W, 128B swizzle; a row
//               of 128 B is 64 fp16 or 128 e4m3 along K) into a 4-stage ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma M=128 N=256 (K=16 fp16 / K=32 e4m3: 32 B of each row either way)
//               x4 per stage; tcgen05.commit frees the stage and, after the last k-block of a chunk, publishes the accumulator
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators)
//   warps 4-11  epilogue (224 registers each via setmaxnreg; the control warps drop to 56): tcgen05.ld chunk -> register
//               accumulators; after the last chunk bias / erf-GELU / relu^2 / rotary, then fp16 hi [+ fp16 lo | + e4m3 planes
//               for the next GEMM] staged in shared memory and written with TMA bulk tensor stores, or TMA reduce-add into the
//               fp32 residual stream
// Roofline: tensor bound. Algorithmic FLOPs = 2*M*N*K; issued tensor-pipe work = x1 / x2 / x3 by mode.
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "common.h"
#include "ptx.cuh"

namespace pg {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UK = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int GEMM_THREADS = (FIRST_EPI_WARP + NUM_EPI_WARPS) * 32;
constexpr uint32_t A_BYTES = BM * BK * 2;  // 16 KiB
constexpr uint32_t B_BYTES = BN * BK * 2;  // 32 KiB
constexpr uint32_t OFF_B = STAGES * A_BYTES;
constexpr uint32_t OFF_STG = OFF_B + STAGES * B_BYTES;      // epilogue staging: 4 KiB per epilogue warp
constexpr uint32_t STG_BYTES = 4096;
constexpr uint32_t OFF_BAR = OFF_STG + NUM_EPI_WARPS * STG_BYTES;
constexpr uint32_t GEMM_SMEM = OFF_BAR + 256 + 1024;  // + barriers + 1 KiB alignment slack
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_SEGS = 24;
// CTA-pair mode (cta_group::2): the pair computes a 256 x 256 tile; each CTA loads its 128 rows of A and HALF of the W tile (128 of
// the 256 rows), so a k-block costs 32 KB of L2 -> SM traffic per CTA instead of 48 KB and the ring holds 6 stages instead of 4.
// ncu r02: the single-CTA kernel pulls 15-17 TB/s through the L2 -> SM crossbar at 75-86 % tensor-pipe activity, i.e. it is bound
// by that traffic, not by the tensor pipe.
constexpr int STAGES2 = 6;
constexpr uint32_t B2_BYTES = 128 * BK * 2;  // 16 KiB: this CTA's half of the W tile
constexpr uint32_t OFF_B2 = STAGES2 * A_BYTES;
static_assert(OFF_B2 + STAGES2 * B2_BYTES == OFF_STG, "both modes use the same 192 KiB of operand stages");

// One run of k-blocks with fixed operand planes. kind: 0 = fp16 (columns in elements), 1 = e4m3 (columns in bytes).
// commit: 1 = the chunk ends after this segment (publish the accumulator, switch TMEM buffer).
struct GemmSeg {
  int a_col, b_col, nkb;
  short kind, commit;
};

struct GemmKParams {
  int M, N, K;
  int nsegs, nchunks;
  int scale_first;        // chunk 0 = e4m3 cross terms: acc = v * a_inv * w_inv[col]
  float a_inv;            // 1 / (2^11 * s_A)
  const float* w_inv;     // [N] 1 / t_n
  int w_uniform;          // every w_inv[n] is the same value (weights packed by the library: one scale per matrix)
  GemmSeg seg[MAX_SEGS];
  const float* bias;
  int out_fmt;            // 0 fp16 hi; 1 fp16 hi + fp16 lo; 2 fp16 hi + e4m3 [lo8 | hi8] byte planes
  float out_scale;        // out_fmt 2: s of the consuming GEMM (hi8 = e4m3(hi*s), lo8 = e4m3(lo*2^11*s))
  const float* rot_cos; const float* rot_sin; int rot_T; int rot_dim;
  int tiles_m, tiles_n;
  int prefetch;           // k-blocks of L2 look-ahead for the A operand (0 = off)
  // delta-operand mode (common.h GemmLaunch): shared base rows added before / subtracted after the activation, masked rows skipped
  const float* base_pre; const __half* base_post; int base_T; const int* mask_pos;
  int grp_rows_a, grp_rows_b;  // grouped (block-diagonal) mode: A rows [g*grp_rows_a, (g+1)*grp_rows_a) pair with W rows g*grp_rows_b + n
};

// Exact-erf GELU (esm/modules.py:17-24): 0.5*x*(1+erf(x/sqrt2)) = 0.5*x + 0.5*|x|*erf(|x|/sqrt2).
// erf via Abramowitz & Stegun 7.1.26 (abs error <= 1.5e-7; measured |gelu error| <= 4.7e-7, below torch's own fp32 gelu),
// branch-free: MUFU.RCP + MUFU.EX2 + FMA-pipe ops instead of erff()'s two divergent code paths, evaluated on TWO values per
// instruction (sm_100 FFMA2 / FMUL2: one issue slot for both lanes; the polynomial is carried negated so that 1 - poly*e is a single
// FFMA2): ~9 issue slots per element.
#define PG_F2C(v) f2_pack((v), (v))
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t x = f2_pack(x0, x1), ax = f2_pack(fabsf(x0), fabsf(x1));
  float u0, u1, t0, t1;
  f2_unpack(f2_fma(ax, PG_F2C(0.23164189f), PG_F2C(1.0f)), u0, u1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const uint64_t t = f2_pack(t0, t1);
  uint64_t npoly = f2_fma(PG_F2C(-1.061405429f), t, PG_F2C(1.453152027f));
  npoly = f2_fma(npoly, t, PG_F2C(-1.421413741f));
  npoly = f2_fma(npoly, t, PG_F2C(0.284496736f));
  npoly = f2_fma(npoly, t, PG_F2C(-0.254829592f));
  npoly = f2_mul(npoly, t);
  float s0, s1, e0, e1;
  f2_unpack(f2_mul(f2_mul(x, x), PG_F2C(-0.72134752044448170368f)), s0, s1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
  const uint64_t erf_abs = f2_fma(npoly, f2_pack(e0, e1), PG_F2C(1.0f));
  f2_unpack(f2_fma(f2_mul(ax, PG_F2C(0.5f)), erf_abs, f2_mul(x, PG_F2C(0.5f))), x0, x1);
}

// Epilogue staging: each warp owns a 4 KiB tile in shared memory laid out for a swizzled TMA store, so the per-thread row writes
// are bank-conflict free and the global write is one coalesced bulk tensor store (or reduce-add) per block, clipped at M and N.
//   fp16 / fp32: 32 rows x 128 B, SWIZZLE_128B (16-byte chunk c of row r lives at chunk c ^ (r & 7))
//   e4m3       : two tiles of 32 rows x 64 B (lo8 at +0, hi8 at +2048), SWIZZLE_64B (chunk c of row r at c ^ ((r >> 1) & 3))
template <int OFF>
__device__ __forceinline__ void stage_row_f16(uint8_t* stg, int lane, const float (&acc)[128], bool lo_plane) {
  uint8_t* row = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float x0 = acc[OFF + c * 8 + 2 * u], x1 = acc[OFF + c * 8 + 2 * u + 1];
      const uint32_t hi = cvt_f16x2_rn(x0, x1);
      if (lo_plane) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        float q0, q1;
        f2_unpack(f2_sub(f2_pack(x0, x1), f2_pack(hf.x, hf.y)), q0, q1);
        w[u] = cvt_f16x2_rn(q0, q1);
      } else {
        w[u] = hi;
      }
    }
    *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// e4m3 planes of 64 accumulated columns as packed words (16 x lo8, 16 x hi8): pure register math, so it can run while the TMA
// store of the fp16 hi tile is still reading the staging buffer.
template <int OFF>
__device__ __forceinline__ void pack_row_f8(const float (&acc)[128], float s_hi, float s_lo, uint32_t (&wl)[16], uint32_t (&wh)[16]) {
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const float x0 = acc[OFF + 4 * u], x1 = acc[OFF + 4 * u + 1], x2 = acc[OFF + 4 * u + 2], x3 = acc[OFF + 4 * u + 3];
    const uint32_t h01 = cvt_f16x2_rn(x0, x1), h23 = cvt_f16x2_rn(x2, x3);
    const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
    const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
    const uint64_t p01 = f2_pack(f01.x, f01.y), p23 = f2_pack(f23.x, f23.y), sh2 = f2_pack(s_hi, s_hi), sl2 = f2_pack(s_lo, s_lo);
    float a0, a1, a2, a3, b0, b1, b2, b3;
    f2_unpack(f2_mul(p01, sh2), a0, a1);
    f2_unpack(f2_mul(p23, sh2), a2, a3);
    f2_unpack(f2_mul(f2_sub(f2_pack(x0, x1), p01), sl2), b0, b1);
    f2_unpack(f2_mul(f2_sub(f2_pack(x2, x3), p23), sl2), b2, b3);
    wh[u] = pack4_e4m3(a0, a1, a2, a3);
    wl[u] = pack4_e4m3(b0, b1, b2, b3);
  }
}
__device__ __forceinline__ void stage_row_f8(uint8_t* stg, int lane, const uint32_t (&wl)[16], const uint32_t (&wh)[16]) {
  uint8_t* rlo = stg + lane * 64;
  uint8_t* rhi = stg + 2048 + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<uint4*>(rlo + ((c ^ sw) << 4)) = make_uint4(wl[4 * c], wl[4 * c + 1], wl[4 * c + 2], wl[4 * c + 3]);
    *reinterpret_cast<uint4*>(rhi + ((c ^ sw) << 4)) = make_uint4(wh[4 * c], wh[4 * c + 1], wh[4 * c + 2], wh[4 * c + 3]);
  }
}
template <int OFF>
__device__ __forceinline__ void stage_row_f32(uint8_t* stg, int lane, const float (&acc)[128]) {
  uint8_t* row = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(row + ((c ^ sw) << 4)) =
        make_float4(acc[OFF + 4 * c], acc[OFF + 4 * c + 1], acc[OFF + 4 * c + 2], acc[OFF + 4 * c + 3]);
}

struct EpiCtx {
  const CUtensorMap* tmHi; const CUtensorMap* tmLo; const CUtensorMap* tmRes;
  uint8_t* stg;
  int lane, gcol, grow0;
  long long row; bool row_ok;
};

// Bias, activation and store of one 64-column group (G = 0, 1) of this thread's 128 accumulated columns.
template <int EPI, int G, int DELTA>
__device__ __forceinline__ void finalize_group(float (&acc)[128], const GemmKParams& p, const EpiCtx& c) {
  constexpr int O = G * 64;
  const int gcol = c.gcol + O;
  if (gcol >= p.N) return;  // warp-uniform
  if (p.bias != nullptr) {
    if (gcol + 64 <= p.N) {  // fast path: 16 vector loads of the (warp-uniform) bias slice
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + gcol);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 x = __ldg(b4 + j);
        f2_unpack(f2_add(f2_pack(acc[O + 4 * j], acc[O + 4 * j + 1]), f2_pack(x.x, x.y)), acc[O + 4 * j], acc[O + 4 * j + 1]);
        f2_unpack(f2_add(f2_pack(acc[O + 4 * j + 2], acc[O + 4 * j + 3]), f2_pack(x.z, x.w)), acc[O + 4 * j + 2], acc[O + 4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[O + j] += (gcol + j < p.N) ? __ldg(p.bias + gcol + j) : 0.f;
    }
  }
  if (DELTA && p.base_pre != nullptr) {
    // delta-operand mode: the shared base row (W a0[t] + b) joins the accumulator before the activation. Read here, 64 columns at a
    // time; starting the accumulators from it at the top of the tile (loads in flight while the MMAs run) measured slower for fc1
    // (193 vs 157 ms/step: it forces the two-step accumulator read-out) and neutral for the other three GEMMs
    const float4* b4 = reinterpret_cast<const float4*>(p.base_pre + (c.row % p.base_T) * static_cast<long long>(p.N) + gcol);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 x = __ldg(b4 + j);
      f2_unpack(f2_add(f2_pack(acc[O + 4 * j], acc[O + 4 * j + 1]), f2_pack(x.x, x.y)), acc[O + 4 * j], acc[O + 4 * j + 1]);
      f2_unpack(f2_add(f2_pack(acc[O + 4 * j + 2], acc[O + 4 * j + 3]), f2_pack(x.z, x.w)), acc[O + 4 * j + 2], acc[O + 4 * j + 3]);
    }
  }
  if (EPI == 1 && DELTA && p.base_post != nullptr) {
    // the output is again a difference, GELU(acc) - base_post[t]. base_post is an fp16 plane (any fixed reference works: the next
    // layer's base_pre was computed from these very values), so the 64 values of this group are 8 vector loads = 32 registers, few
    // enough to be requested BEFORE the GELU arithmetic and consumed after it: one of the two L2 round trips per group disappears
    // behind ~600 instructions (with fp32 values, 64 registers, the same reordering had to be split in halves and measured slower)
    const uint4* b4 = reinterpret_cast<const uint4*>(p.base_post + (c.row % p.base_T) * static_cast<long long>(p.N) + gcol);
    uint4 post[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) post[j] = __ldg(b4 + j);
#pragma unroll
    for (int j = 0; j < 64; j += 2) gelu_erf2(acc[O + j], acc[O + j + 1]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t w[4] = {post[j].x, post[j].y, post[j].z, post[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 q = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        const int e = O + 8 * j + 2 * k;
        f2_unpack(f2_sub(f2_pack(acc[e], acc[e + 1]), f2_pack(q.x, q.y)), acc[e], acc[e + 1]);
      }
    }
  } else if (EPI == 1) {
#pragma unroll
    for (int j = 0; j < 64; j += 2) gelu_erf2(acc[O + j], acc[O + j + 1]);
  } else if (EPI == 4) {  // squared ReLU (Tranception MLP, tranception/activations.py:79-84)
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      const float a = fmaxf(acc[O + j], 0.f), b = fmaxf(acc[O + j + 1], 0.f);
      const uint64_t ab = f2_pack(a, b);
      f2_unpack(f2_mul(ab, ab), acc[O + j], acc[O + j + 1]);
    }
  } else if (EPI == 3 && gcol < 2 * p.rot_dim) {
    // rotary: x*cos + rotate_half(x)*sin over one 64-wide head; cos/sin[t, j] for j in [0,32) (both halves equal)
    const int t = static_cast<int>(c.row % p.rot_T);
    const float* cs = p.rot_cos + static_cast<long long>(t) * 32;
    const float* sn = p.rot_sin + static_cast<long long>(t) * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float co = c.row_ok ? __ldg(cs + j) : 1.f, s = c.row_ok ? __ldg(sn + j) : 0.f;
      const float a = acc[O + j], b = acc[O + 32 + j];
      acc[O + j] = a * co - b * s;
      acc[O + 32 + j] = b * co + a * s;
    }
  }
  if (EPI == 2 && DELTA && p.mask_pos != nullptr && c.row_ok) {  // the masked row of each copy is updated by the compact exact path instead
    const long long copy = c.row / p.base_T;
    if (c.row - copy * p.base_T == __ldg(p.mask_pos + copy)) {
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[O + j] = 0.f;
    }
  }
  uint8_t* stg = c.stg;
  const int lane = c.lane;
  if (EPI == 2) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (gcol + hh * 32 >= p.N) break;
      if (lane == 0) bulk_wait_read0();  // the previous bulk store has finished reading this warp's staging tile
      __syncwarp();
      if (hh == 0) stage_row_f32<O>(stg, lane, acc); else stage_row_f32<O + 32>(stg, lane, acc);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(c.tmRes, stg, gcol + hh * 32, c.grow0);
        bulk_commit();
      }
    }
  } else {
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    stage_row_f16<O>(stg, lane, acc, false);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(c.tmHi, stg, gcol, c.grow0);
      bulk_commit();
    }
    if (p.out_fmt == 1) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      stage_row_f16<O>(stg, lane, acc, true);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(c.tmLo, stg, gcol, c.grow0);
        bulk_commit();
      }
    } else if (p.out_fmt == 2) {
      uint32_t wl[16], wh[16];
      pack_row_f8<O>(acc, p.out_scale, p.out_scale * 2048.f, wl, wh);  // overlaps the hi store's read of the staging tile
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      stage_row_f8(stg, lane, wl, wh);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(c.tmLo, stg, gcol, c.grow0);                 // lo8 plane: bytes [0, N) of the e4m3 region
        tma_store_2d(c.tmLo, stg + 2048, p.N + gcol, c.grow0);    // hi8 plane: bytes [N, 2N)
        bulk_commit();
      }
    }
  }
}

template <int EPI, int CTA2, int DELTA>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8,
               const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo,
               const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  constexpr int NST = CTA2 ? STAGES2 : STAGES;                 // stages of the operand ring
  constexpr uint32_t BST = CTA2 ? B2_BYTES : B_BYTES;          // bytes of W this CTA stages per k-block
  uint8_t* smA = smem;
  uint8_t* smB = smem + (CTA2 ? OFF_B2 : OFF_B);
  uint8_t* smStg = smem + OFF_STG;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0;            // 0 = leader of the pair: issues the MMAs
  const int ntiles = p.tiles_m * p.tiles_n;                    // CTA2: tiles_m counts 256-row tiles
  const int tile0 = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tstep = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int row_base = CTA2 ? 2 * BM : BM;                       // rows of the output tile per (pair of) CTA(s)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], CTA2 ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);  // pair mode: the epilogue warps of BOTH CTAs release the leader
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CTA2) { tmem_alloc_2sm(tmem_ptr, TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < FIRST_EPI_WARP) {
    setmaxnreg_dec<56>();
    if (warp == 0) {
      // TMA producer: warp-uniform loop, one elected lane arms the barrier and issues the two bulk tensor loads of a k-block.
      int stage = 0;
      uint32_t phase = 0;
      // Optional L2 look-ahead cursor (p.prefetch k-blocks ahead of the loads, A operand only). Measured neutral-to-negative on the
      // in-model shapes (bench r02: 737 vs 762 ms/step with it off), so the default is off.
      int pf_tile = tile0, pf_s = 0, pf_kb = 0;
      auto pf_step = [&]() {
        if (pf_tile >= ntiles) return;
        const GemmSeg sg = p.seg[pf_s];
        if (elect_one())
          tma_prefetch_l2_2d(sg.kind ? &tmA8 : &tmA, sg.a_col + pf_kb * (sg.kind ? 2 * BK : BK),
                             (pf_tile / p.tiles_n) * row_base + static_cast<int>(rank) * BM);
        __syncwarp();
        if (++pf_kb == sg.nkb) {
          pf_kb = 0;
          if (++pf_s == p.nsegs) { pf_s = 0; pf_tile += tstep; }
        }
      };
      for (int i = 0; i < p.prefetch; ++i) pf_step();
      for (int tile = tile0; tile < ntiles; tile += tstep) {
        const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
        const int a_row = m_blk * row_base + static_cast<int>(rank) * BM;          // this CTA's 128 rows of A
        int b_row = n_blk * BN + (CTA2 ? static_cast<int>(rank) * 128 : 0);         // pair mode: this CTA's half of the W tile
        if (p.grp_rows_a) b_row += (m_blk * row_base / p.grp_rows_a) * p.grp_rows_b;  // grouped: this group's block of W rows
        for (int s = 0; s < p.nsegs; ++s) {
          const GemmSeg sg = p.seg[s];
          const CUtensorMap* ma = sg.kind ? &tmA8 : &tmA;
          const CUtensorMap* mb = sg.kind ? &tmB8 : &tmB;
          const int step = sg.kind ? 2 * BK : BK;  // 128 B of a row: 64 fp16 or 128 e4m3
          for (int kb = 0; kb < sg.nkb; ++kb) {
            if (p.prefetch) pf_step();
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              if (CTA2) {
                // both CTAs' bytes complete on the LEADER's barrier (its MMA thread is the only consumer); the leader arms it
                if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + B2_BYTES));
                tma_load_2d_2sm(smA + stage * A_BYTES, ma, &full[stage], sg.a_col + kb * step, a_row);
                tma_load_2d_2sm(smB + stage * B2_BYTES, mb, &full[stage], sg.b_col + kb * step, b_row);
              } else {
                mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
                tma_load_2d(smA + stage * A_BYTES, ma, &full[stage], sg.a_col + kb * step, a_row);
                tma_load_2d(smB + stage * B_BYTES, mb, &full[stage], sg.b_col + kb * step, b_row);
              }
            }
            __syncwarp();
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      // All 32 lanes run the (warp-uniform) control flow and the barrier waits; one elected lane issues the MMAs and commits.
      // Pair mode: only the leader CTA issues; the instruction reads A rows and W rows from both CTAs' shared memory (same offsets)
      // and accumulates 128 rows in each CTA's TMEM; commits are multicast to the barriers of both CTAs.
      constexpr uint32_t idesc = make_idesc_f16(CTA2 ? 2 * BM : BM, BN, 0, 0);  // same bit pattern for kind::f8f6f4 with e4m3 operands
      const uint32_t a_base = smem_u32(smA), b_base = smem_u32(smB);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t g = 0;  // running chunk number: TMEM buffer g & 1, barrier phase (g >> 1) & 1
      for (int tile = tile0; tile < ntiles; tile += tstep) {
        bool fresh = true;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nsegs; ++s) {
          const GemmSeg sg = p.seg[s];
          const uint32_t buf = g & 1;
          if (fresh) {
            mbar_wait(&tempty[buf], ((g >> 1) & 1) ^ 1);
            tc_fence_after();
            accumulate = 0;
            fresh = false;
          }
          const uint32_t tmem_d = tmem_base + buf * BN;
          for (int kb = 0; kb < sg.nkb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = make_desc_sw128(a_base + stage * A_BYTES, 1024);
            const uint64_t bdesc = make_desc_sw128(b_base + stage * BST, 1024);
            if (elect_one()) {
              // 4 x (K = 32 bytes of every row): 16 fp16 or 32 e4m3 per instruction
              if (sg.kind) umma_kblock<1, CTA2 ? 2 : 1>(tmem_d, adesc, bdesc, idesc, accumulate);
              else umma_kblock<0, CTA2 ? 2 : 1>(tmem_d, adesc, bdesc, idesc, accumulate);
              if (CTA2) umma_commit_2sm(&empty[stage]); else umma_commit(&empty[stage]);
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
          if (sg.commit) {
            if (elect_one()) { if (CTA2) umma_commit_2sm(&tfull[buf]); else umma_commit(&tfull[buf]); }
            __syncwarp();
            ++g;
            fresh = true;
          }
        }
      }
    }
  } else {
    setmaxnreg_inc<224>();
    const int q = warp & 3;                           // TMEM lane quadrant this warp may read
    const int half_id = (warp - FIRST_EPI_WARP) >> 2;  // which 128-column half of the tile
    uint32_t g = 0;
    EpiCtx c;
    c.tmHi = &tmHi; c.tmLo = &tmLo; c.tmRes = &tmRes;
    c.stg = smStg + (warp - FIRST_EPI_WARP) * STG_BYTES;
    c.lane = lane;
    for (int tile = tile0; tile < ntiles; tile += tstep) {
      const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
      c.gcol = n_blk * BN + half_id * 128;
      c.grow0 = m_blk * row_base + static_cast<int>(rank) * BM + q * 32;
      c.row = static_cast<long long>(c.grow0) + lane;
      c.row_ok = c.row < p.M;
      float acc[128];
      for (int ch = 0; ch < p.nchunks; ++ch, ++g) {
        const uint32_t buf = g & 1;
        mbar_wait(&tfull[buf], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + half_id * 128;
        if (ch == 0) {
          // first chunk: nothing accumulated yet, so all four 32-column loads go out together straight into the accumulators
          uint32_t r[4][32];
#pragma unroll
          for (int i = 0; i < 4; ++i) tmem_ld_32x32b_x32(taddr + i * 32, r[i]);
          tmem_ld_wait();
          if (!p.scale_first) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[i * 32 + j] = __uint_as_float(r[i][j]);
          } else if (p.w_uniform) {  // e4m3 cross terms, one scale per weight matrix: a single factor for the whole tile
            const float f = p.a_inv * __ldg(p.w_inv);
            // scalar on purpose (as the chunk adds below): with 128 accumulators + 128 freshly loaded values live, the 64-bit
            // register pairing of the packed forms made ptxas spill
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[i * 32 + j] = __uint_as_float(r[i][j]) * f;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = c.gcol + i * 32 + j;
                acc[i * 32 + j] = __uint_as_float(r[i][j]) * ((col < p.N) ? p.a_inv * __ldg(p.w_inv + col) : 0.f);
              }
          }
        } else {
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(taddr + hp * 64, r0);
            tmem_ld_32x32b_x32(taddr + hp * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc[hp * 64 + j] += __uint_as_float(r0[j]);
              acc[hp * 64 + 32 + j] += __uint_as_float(r1[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2 && rank != 0) mbar_arrive_cluster(&tempty[buf], 0);  // the leader's MMA thread waits on its own barrier
          else mbar_arrive(&tempty[buf]);
        }
      }
      finalize_group<EPI, 0, DELTA>(acc, p, c);
      finalize_group<EPI, 1, DELTA>(acc, p, c);
    }
    if (lane == 0) bulk_wait0();  // all bulk stores of this warp have landed
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // neither CTA's shared / tensor memory may go away while the pair still uses it
  if (warp == 2) {
    if (CTA2) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Host-side cache of encoded tensor maps: the model issues the same few (buffer, shape) combinations thousands of times per
// assay. thread_local: one map per calling thread, so handles driven from different threads never share state.
struct TmapKey {
  const void* ptr; uint64_t rows, cols, ld; uint32_t box_rows, box_cols; int elem_bytes, swizzle;
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");

}  // namespace

// 2D row-major tensor [rows, cols] with row pitch ld (elements); box = [box_rows, box_cols], out-of-bounds reads give zero,
// out-of-bounds parts of a store are clipped. elem_bytes 2 = fp16, 4 = fp32, 1 = bytes (e4m3 planes). swizzle in bytes: 128 or 64.
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                 int elem_bytes, int swizzle) {
  TmapKey key;
  std::memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rows = rows; key.cols = cols; key.ld = ld; key.box_rows = box_rows; key.box_cols = box_cols;
  key.elem_bytes = elem_bytes; key.swizzle = swizzle;
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  auto it = cache.find(key);
  if (it != cache.end()) {
    *m = it->second;
    return PG_OK;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(PG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * elem_bytes) % 16)
    return set_error(PG_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                 : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = fn(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PG_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(static_cast<int>(r)));
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *m);
  return PG_OK;
}
int make_tmap_f16_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  return make_tmap_2d(m, ptr, rows, cols, ld, box_rows, box_cols, 2, 128);
}

// Longest run of K one hi*hi chunk accumulates before the epilogue takes it over (see the header). Default 1280; the environment
// variable PG_GEMM_KCHUNK or pg_set_tuning("gemm_kchunk", v) override it; 0 = no chunking (round-1 behaviour, for the probe).
static int g_kchunk = -1, g_prefetch = -1, g_cta2 = -1;
int gemm_cta2() {
  if (g_cta2 < 0) {
    const char* e = getenv("PG_GEMM_CTA2");
    g_cta2 = e ? atoi(e) : 1;  // default: CTA-pair kernel (bench r02: 690 vs 720 ms/step, issued 0.85 vs 0.79 of the sustained peak)
  }
  return g_cta2;
}
void set_gemm_cta2(int v) { g_cta2 = v ? 1 : 0; }
int gemm_kchunk() {
  if (g_kchunk < 0) {
    const char* e = getenv("PG_GEMM_KCHUNK");
    g_kchunk = e ? atoi(e) : 1280;
  }
  return g_kchunk < BK ? (1 << 30) : g_kchunk;
}
void set_gemm_kchunk(int v) { g_kchunk = v < 0 ? 0 : v; }
int gemm_prefetch() {
  if (g_prefetch < 0) {
    const char* e = getenv("PG_GEMM_PREFETCH");
    g_prefetch = e ? atoi(e) : 0;
  }
  return g_prefetch;
}
void set_gemm_prefetch(int v) { g_prefetch = v < 0 ? 0 : (v > 64 ? 64 : v); }

int launch_gemm(const GemmLaunch& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return set_error(PG_ERR_ARG, "gemm: empty problem");
  if (g.K % BK) return set_error(PG_ERR_ARG, "gemm: K must be a multiple of 64");
  if (g.nseg < 1 || g.nseg > 3) return set_error(PG_ERR_ARG, "gemm: nseg must be 1, 2 (fp16 + e4m3 cross terms) or 3");
  if (g.epi < 0 || g.epi > 4) return set_error(PG_ERR_ARG, "gemm: bad epilogue");
  if (g.epi == 2 ? !g.resid : !g.out) return set_error(PG_ERR_ARG, "gemm: missing output");
  if (g.epi == 3 && (!g.rot_cos || !g.rot_sin || g.rot_T <= 0 || g.rot_dim % 64)) return set_error(PG_ERR_ARG, "gemm: bad rotary args");
  if (g.nseg == 2 && (!g.w_inv || !(g.a_scale > 0.f))) return set_error(PG_ERR_ARG, "gemm: nseg 2 needs w_inv[N] and a_scale > 0");
  if (g.out_fmt < 0 || g.out_fmt > 2) return set_error(PG_ERR_ARG, "gemm: bad out_fmt");
  if ((g.base_pre || g.base_post || g.mask_pos) && (!g.base_pre || g.base_T <= 0 || g.N % 64 || g.M % g.base_T || g.grp_rows_a ||
                                                       (g.mask_pos && g.epi != 2) || (g.base_post && g.epi == 2)))
    return set_error(PG_ERR_ARG, "gemm: delta-operand mode needs base_pre, base_T > 0 dividing M, N % 64 == 0, no groups; mask_pos only with the residual epilogue");
  if (g.epi != 2 && g.out_fmt == 2 && (g.N % 64 || !(g.out_scale > 0.f)))
    return set_error(PG_ERR_ARG, "gemm: out_fmt 2 needs N % 64 == 0 and out_scale > 0");
  if (g.epi != 2 && g.out_fmt >= 1 && g.out_lo_off <= 0) return set_error(PG_ERR_ARG, "gemm: out_fmt 1/2 need out_lo_off > 0");
  int dev = 0;
  PG_CUDA_OK(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (dev < 64 && !attr_set[dev]) {
#define PG_SET_SMEM(E)                                                                                              \
  PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<E, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM)); \
  PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<E, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM))
    PG_SET_SMEM(0); PG_SET_SMEM(1); PG_SET_SMEM(2); PG_SET_SMEM(3); PG_SET_SMEM(4);
#undef PG_SET_SMEM
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    attr_set[dev] = true;
  }
  const uint64_t K = static_cast<uint64_t>(g.K);
  const uint64_t width = K * (g.nseg == 3 ? 2 : 1);
  const bool grouped = g.grp_rows_a > 0;
  // CTA pairs work on 256-row tiles; a grouped problem whose groups are only a multiple of 128 rows runs the single-CTA form
  const int cta2 = gemm_cta2() && !(grouped && g.grp_rows_a % (2 * BM));
  const uint32_t wbox = cta2 ? 128 : BN;  // rows of W one CTA stages per k-block (pair mode: its half of the 256-row tile)
  CUtensorMap tmA, tmB, tmA8{}, tmB8{};
  int rc = make_tmap_2d(&tmA, g.a, g.M, width, g.lda, BM, BK, 2, 128);
  if (rc) return rc;
  if (grouped && (g.nseg == 2 || g.grp_rows_a % BM || g.grp_rows_b < g.N || g.M % g.grp_rows_a || g.bias || g.epi == 3))
    return set_error(PG_ERR_ARG, "gemm: grouped mode needs fp16 operands, A groups of a multiple of 128 rows, W groups of >= N rows, no bias");
  const uint64_t w_rows = grouped ? static_cast<uint64_t>(g.M / g.grp_rows_a) * g.grp_rows_b : static_cast<uint64_t>(g.N);
  rc = make_tmap_2d(&tmB, g.w, w_rows, width, g.ldw, wbox, BK, 2, 128);
  if (rc) return rc;
  if (g.nseg == 2) {  // e4m3 planes follow the fp16 hi plane of each row: bytes [2K, 4K) = K-concatenated [lo8 | hi8] / [hi8 | lo8]
    rc = make_tmap_2d(&tmA8, static_cast<const uint8_t*>(g.a) + 2 * K, g.M, 2 * K, static_cast<uint64_t>(g.lda) * 2, BM, 2 * BK, 1, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB8, static_cast<const uint8_t*>(g.w) + 2 * K, g.N, 2 * K, static_cast<uint64_t>(g.ldw) * 2, wbox, 2 * BK, 1, 128);
    if (rc) return rc;
  }
  CUtensorMap tmHi{}, tmLo{}, tmRes{};
  if (g.epi == 2) {
    rc = make_tmap_2d(&tmRes, g.resid, g.M, g.N, g.ldr, 32, 32, 4, 128);
    if (rc) return rc;
  } else {
    rc = make_tmap_2d(&tmHi, g.out, g.M, g.N, g.ldo, 32, 64, 2, 128);
    if (rc) return rc;
    if (g.out_fmt == 1) {
      rc = make_tmap_2d(&tmLo, g.out + g.out_lo_off, g.M, g.N, g.ldo, 32, 64, 2, 128);
      if (rc) return rc;
    } else if (g.out_fmt == 2) {
      rc = make_tmap_2d(&tmLo, g.out + g.out_lo_off, g.M, 2 * static_cast<uint64_t>(g.N), static_cast<uint64_t>(g.ldo) * 2, 32, 64, 1, 64);
      if (rc) return rc;
    }
  }
  GemmKParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  const int kblocks = g.K / BK;
  int ns = 0, nchunks = 0;
  auto push = [&](int a_col, int b_col, int nkb, int kind, int commit) {
    p.seg[ns].a_col = a_col; p.seg[ns].b_col = b_col; p.seg[ns].nkb = nkb;
    p.seg[ns].kind = static_cast<short>(kind); p.seg[ns].commit = static_cast<short>(commit);
    ++ns;
    nchunks += commit;
  };
  // hi*hi pieces: as equal as possible, each <= kchunk along K, at most what the segment table holds
  int pieces = (g.K + gemm_kchunk() - 1) / gemm_kchunk();
  if (pieces < 1) pieces = 1;
  if (pieces > MAX_SEGS - 2) pieces = MAX_SEGS - 2;
  if (pieces > kblocks) pieces = kblocks;
  if (g.nseg == 2) {
    push(0, 0, kblocks, 1, 1);  // [lo8 | hi8] x [hi8 | lo8]: 2K bytes = kblocks blocks of 128 B
    p.scale_first = 1;
    p.a_inv = 1.0f / (2048.f * g.a_scale);
    p.w_inv = g.w_inv;
    p.w_uniform = g.w_uniform;
  } else if (g.nseg == 3) {
    push(g.K, 0, kblocks, 0, 0);  // lo*hi
    push(0, g.K, kblocks, 0, 1);  // hi*lo
  }
  for (int i = 0, kb0 = 0; i < pieces; ++i) {
    const int nkb = kblocks / pieces + (i < kblocks % pieces ? 1 : 0);
    push(kb0 * BK, kb0 * BK, nkb, 0, 1);
    kb0 += nkb;
  }
  p.nsegs = ns; p.nchunks = nchunks;
  p.bias = g.bias;
  p.out_fmt = (g.epi == 2) ? 0 : g.out_fmt;
  p.out_scale = g.out_scale;
  p.rot_cos = g.rot_cos; p.rot_sin = g.rot_sin; p.rot_T = g.rot_T; p.rot_dim = g.rot_dim;
  p.tiles_m = cta2 ? (g.M + 2 * BM - 1) / (2 * BM) : (g.M + BM - 1) / BM;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.prefetch = grouped ? 0 : gemm_prefetch();
  p.grp_rows_a = g.grp_rows_a; p.grp_rows_b = g.grp_rows_b;
  p.base_pre = g.base_pre; p.base_post = g.base_post; p.base_T = g.base_T; p.mask_pos = g.mask_pos;
  const int ntiles = p.tiles_m * p.tiles_n;
  const bool delta = g.base_pre != nullptr;
  if (delta && (!cta2 || g.epi > 2)) return set_error(PG_ERR_UNSUPPORTED, "gemm: the delta-operand form runs on the CTA-pair kernel with epilogues 0, 1, 2");
  if (!cta2) {
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    switch (g.epi) {
      case 0: gemm_tc_kernel<0, 0, 0><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
      case 1: gemm_tc_kernel<1, 0, 0><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
      case 2: gemm_tc_kernel<2, 0, 0><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
      case 3: gemm_tc_kernel<3, 0, 0><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
      default: gemm_tc_kernel<4, 0, 0><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
    }
    PG_CUDA_OK(cudaGetLastError());
    return PG_OK;
  }
  // pair mode: clusters of two CTAs (same TPC), one cluster per pair of SMs
  const int pairs = ntiles < num_sms() / 2 ? ntiles : num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GEMM_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (delta) {
    switch (g.epi) {
      case 0: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<0, 1, 1>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
      case 1: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1, 1, 1>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
      default: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2, 1, 1>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
    }
    PG_CUDA_OK(e);
    PG_CUDA_OK(cudaGetLastError());
    return PG_OK;
  }
  switch (g.epi) {
    case 0: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<0, 1, 0>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
    case 1: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1, 1, 0>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
    case 2: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2, 1, 0>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
    case 3: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<3, 1, 0>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
    default: e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<4, 1, 0>, tmA, tmB, tmA8, tmB8, tmHi, tmLo, tmRes, p); break;
  }
  PG_CUDA_OK(e);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
