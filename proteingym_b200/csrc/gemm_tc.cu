// K3 — tcgen05/TMA GEMM with fused epilogues for the linear layers of the ESM / Tranception forward
// (reference ops: esm/multihead_attention.py:243-261,395; esm/modules.py:138-140).
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T + bias )        A, W fp16 K-major (row-major), fp32 accumulate in TMEM.
//
// Precision: the reference is fp32. With nseg == 3 each operand is an fp16 (hi | lo) pair stored as two column blocks
// and the K loop runs three segments  hi*hi + lo*hi + hi*lo  into the same TMEM accumulator (~22-bit operands).
//
// Structure (one persistent CTA per SM, 384 threads, no clusters):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D tiles (128x64 of A, 256x64 of W, 128B swizzle) into a 4-stage ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16) x4 per stage; tcgen05.commit frees the
//               stage and, after the last k-block, publishes the accumulator
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, so the epilogue of tile i overlaps the
//               MMAs of tile i+1)
//   warps 4-11  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> bias / erf-GELU / residual / rotary ->
//               global (fp16 hi[/lo] or fp32 residual stream)
// Roofline: tensor bound. Algorithmic FLOPs = 2*M*N*K*nseg.
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

struct GemmKParams {
  int M, N, K, nseg;
  int a_off[3], b_off[3];
  const float* bias;
  int epi;
  __half* out; long long ldo; long long lo_off;
  float* resid; long long ldr;
  const float* rot_cos; const float* rot_sin; int rot_T; int rot_dim;
  int tiles_m, tiles_n;
};

// Exact-erf GELU (esm/modules.py:17-24): 0.5*x*(1+erf(x/sqrt2)) = 0.5*x + 0.5*|x|*erf(|x|/sqrt2).
// erf via Abramowitz & Stegun 7.1.26 (abs error <= 1.5e-7; measured |gelu error| <= 4.7e-7, below torch's own fp32 gelu),
// branch-free: MUFU.RCP + MUFU.EX2 + ~11 FMA-pipe ops instead of erff()'s two divergent code paths, so the fc1 epilogue
// fits under the MMAs of the next tile even in single-pass mode.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, ax, 1.0f)));  // 1/(1 + 0.3275911*|x|/sqrt2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));  // exp(-x^2/2)
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return fmaf(0.5f * ax, erf_abs, 0.5f * x);
}

// Epilogue staging: each warp owns a 32-row x 128-byte tile in shared memory laid out for a SWIZZLE_128B TMA store
// (16-byte chunk c of row r lives at chunk c ^ (r & 7)), so the per-thread row writes are bank-conflict free and the
// global write is one coalesced bulk tensor store (or reduce-add) per 32x64 fp16 / 32x32 fp32 block, clipped at M and N.
__device__ __forceinline__ void stage_row_f16(uint8_t* stg, int lane, const float (&v0)[32], const float (&v1)[32], bool lo_plane) {
  uint8_t* row = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float x0 = (c < 4) ? v0[c * 8 + 2 * u] : v1[(c - 4) * 8 + 2 * u];
      const float x1 = (c < 4) ? v0[c * 8 + 2 * u + 1] : v1[(c - 4) * 8 + 2 * u + 1];
      const uint32_t hi = cvt_f16x2_rn(x0, x1);
      if (lo_plane) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        w[u] = cvt_f16x2_rn(x0 - hf.x, x1 - hf.y);
      } else {
        w[u] = hi;
      }
    }
    *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ void stage_row_f32(uint8_t* stg, int lane, const float (&v)[32]) {
  uint8_t* row = stg + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(row + ((c ^ sw) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo,
               const __grid_constant__ CUtensorMap tmRes, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint8_t* smA = smem;
  uint8_t* smB = smem + OFF_B;
  uint8_t* smStg = smem + OFF_STG;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = p.K / BK;
  const int ntiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
        for (int seg = 0; seg < p.nseg; ++seg) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
            tma_load_2d(smA + stage * A_BYTES, &tmA, &full[stage], p.a_off[seg] + kb * BK, m_blk * BM);
            tma_load_2d(smB + stage * B_BYTES, &tmB, &full[stage], p.b_off[seg] + kb * BK, n_blk * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        mbar_wait(&tempty[buf], tphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        uint32_t accumulate = 0;
        for (int seg = 0; seg < p.nseg; ++seg) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = make_desc_sw128(smem_u32(smA + stage * A_BYTES), 1024);
            const uint64_t bdesc = make_desc_sw128(smem_u32(smB + stage * B_BYTES), 1024);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              // advance 16 fp16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
              umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else if (warp >= FIRST_EPI_WARP) {
    const int q = warp & 3;                           // TMEM lane quadrant this warp may read
    const int half_id = (warp - FIRST_EPI_WARP) >> 2;  // which 128-column half of the tile
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / p.tiles_n, n_blk = tile % p.tiles_n;
      const int buf = it & 1;
      const uint32_t tphase = (it >> 1) & 1;
      mbar_wait(&tfull[buf], tphase);
      tc_fence_after();
      const long long row = static_cast<long long>(m_blk) * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {  // pairs of 32-column chunks = one 64-wide head
        const int col0 = half_id * 128 + cp * 64;
        const int gcol = n_blk * BN + col0;
        if (gcol >= p.N) break;  // warp-uniform
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + col0;
        tmem_ld_32x32b_x32(taddr, r0);
        tmem_ld_32x32b_x32(taddr + 32, r1);
        tmem_ld_wait();
        float v0[32], v1[32];
        if (p.bias != nullptr && gcol + 64 <= p.N) {  // fast path: 16 vector loads of the (warp-uniform) bias slice
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + gcol);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = __ldg(b4 + j), y = __ldg(b4 + 8 + j);
            v0[4 * j] = __uint_as_float(r0[4 * j]) + x.x; v0[4 * j + 1] = __uint_as_float(r0[4 * j + 1]) + x.y;
            v0[4 * j + 2] = __uint_as_float(r0[4 * j + 2]) + x.z; v0[4 * j + 3] = __uint_as_float(r0[4 * j + 3]) + x.w;
            v1[4 * j] = __uint_as_float(r1[4 * j]) + y.x; v1[4 * j + 1] = __uint_as_float(r1[4 * j + 1]) + y.y;
            v1[4 * j + 2] = __uint_as_float(r1[4 * j + 2]) + y.z; v1[4 * j + 3] = __uint_as_float(r1[4 * j + 3]) + y.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float b0 = (p.bias && gcol + j < p.N) ? __ldg(p.bias + gcol + j) : 0.f;
            const float b1 = (p.bias && gcol + 32 + j < p.N) ? __ldg(p.bias + gcol + 32 + j) : 0.f;
            v0[j] = __uint_as_float(r0[j]) + b0;
            v1[j] = __uint_as_float(r1[j]) + b1;
          }
        }
        if (EPI == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v0[j] = gelu_erf(v0[j]);
            v1[j] = gelu_erf(v1[j]);
          }
        } else if (EPI == 4) {  // squared ReLU (Tranception MLP, tranception/activations.py:79-84)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = fmaxf(v0[j], 0.f), b = fmaxf(v1[j], 0.f);
            v0[j] = a * a;
            v1[j] = b * b;
          }
        } else if (EPI == 3 && gcol < 2 * p.rot_dim) {
          // rotary: x*cos + rotate_half(x)*sin over one 64-wide head; cos/sin[t, j] for j in [0,32) (both halves equal)
          const int t = static_cast<int>(row % p.rot_T);
          const float* cs = p.rot_cos + static_cast<long long>(t) * 32;
          const float* sn = p.rot_sin + static_cast<long long>(t) * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float c = row_ok ? __ldg(cs + j) : 1.f, s = row_ok ? __ldg(sn + j) : 0.f;
            const float a = v0[j], b = v1[j];
            v0[j] = a * c - b * s;
            v1[j] = b * c + a * s;
          }
        }
        uint8_t* stg = smStg + (warp - FIRST_EPI_WARP) * STG_BYTES;
        const int grow0 = m_blk * BM + q * 32;
        if (EPI == 2) {
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            if (gcol + hh * 32 >= p.N) break;
            if (lane == 0) bulk_wait_read0();  // the previous bulk store has finished reading this warp's staging tile
            __syncwarp();
            if (hh == 0) stage_row_f32(stg, lane, v0); else stage_row_f32(stg, lane, v1);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(&tmRes, stg, gcol + hh * 32, grow0);
              bulk_commit();
            }
          }
        } else {
#pragma unroll 1
          for (int pl = 0; pl < 2; ++pl) {
            if (pl == 1 && p.lo_off <= 0) break;
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            stage_row_f16(stg, lane, v0, v1, pl == 1);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (pl) tma_store_2d(&tmLo, stg, gcol, grow0);
              else tma_store_2d(&tmHi, stg, gcol, grow0);
              bulk_commit();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  }

  if (warp >= FIRST_EPI_WARP && lane == 0) bulk_wait0();  // all bulk stores of this warp have landed
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
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

}  // namespace

// 2D fp16 row-major tensor [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle,
// out-of-bounds elements read as zero.
static int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                        uint32_t box_cols, int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(PG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * elem_bytes) % 16)
    return set_error(PG_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PG_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(static_cast<int>(r)));
  return PG_OK;
}
int make_tmap_f16_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  return make_tmap_2d(m, ptr, rows, cols, ld, box_rows, box_cols, 2);
}

int launch_gemm(const GemmLaunch& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return set_error(PG_ERR_ARG, "gemm: empty problem");
  if (g.K % BK) return set_error(PG_ERR_ARG, "gemm: K must be a multiple of 64");
  if (g.nseg != 1 && g.nseg != 3) return set_error(PG_ERR_ARG, "gemm: nseg must be 1 or 3");
  if (g.epi < 0 || g.epi > 4) return set_error(PG_ERR_ARG, "gemm: bad epilogue");
  if (g.epi == 2 ? !g.resid : !g.out) return set_error(PG_ERR_ARG, "gemm: missing output");
  if (g.epi == 3 && (!g.rot_cos || !g.rot_sin || g.rot_T <= 0 || g.rot_dim % 64)) return set_error(PG_ERR_ARG, "gemm: bad rotary args");
  static bool attr_set = false;
  if (!attr_set) {
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    PG_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    attr_set = true;
  }
  const uint64_t width = static_cast<uint64_t>(g.K) * (g.nseg == 3 ? 2 : 1);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_f16_2d(&tmA, g.a, g.M, width, g.lda, BM, BK);
  if (rc) return rc;
  rc = make_tmap_f16_2d(&tmB, g.w, g.N, width, g.ldw, BN, BK);
  if (rc) return rc;
  CUtensorMap tmHi{}, tmLo{}, tmRes{};
  if (g.epi == 2) {
    rc = make_tmap_2d(&tmRes, g.resid, g.M, g.N, g.ldr, 32, 32, 4);
    if (rc) return rc;
  } else {
    rc = make_tmap_2d(&tmHi, g.out, g.M, g.N, g.ldo, 32, 64, 2);
    if (rc) return rc;
    if (g.out_lo_off > 0) {
      rc = make_tmap_2d(&tmLo, g.out + g.out_lo_off, g.M, g.N, g.ldo, 32, 64, 2);
      if (rc) return rc;
    }
  }
  GemmKParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.nseg = g.nseg;
  // segments: hi*hi, lo*hi, hi*lo
  p.a_off[0] = 0; p.b_off[0] = 0;
  p.a_off[1] = g.K; p.b_off[1] = 0;
  p.a_off[2] = 0; p.b_off[2] = g.K;
  p.bias = g.bias; p.epi = g.epi;
  p.out = g.out; p.ldo = g.ldo; p.lo_off = g.out_lo_off;
  p.resid = g.resid; p.ldr = g.ldr;
  p.rot_cos = g.rot_cos; p.rot_sin = g.rot_sin; p.rot_T = g.rot_T; p.rot_dim = g.rot_dim;
  p.tiles_m = (g.M + BM - 1) / BM;
  p.tiles_n = (g.N + BN - 1) / BN;
  const int ntiles = p.tiles_m * p.tiles_n;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  switch (g.epi) {
    case 0: gemm_tc_kernel<0><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmHi, tmLo, tmRes, p); break;
    case 1: gemm_tc_kernel<1><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmHi, tmLo, tmRes, p); break;
    case 2: gemm_tc_kernel<2><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmHi, tmLo, tmRes, p); break;
    case 3: gemm_tc_kernel<3><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmHi, tmLo, tmRes, p); break;
    default: gemm_tc_kernel<4><<<grid, GEMM_THREADS, GEMM_SMEM, s>>>(tmA, tmB, tmHi, tmLo, tmRes, p); break;
  }
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
