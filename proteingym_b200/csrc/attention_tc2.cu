// K4 (v3) — tcgen05/TMEM attention with TWO query tiles in flight per CTA ("ping-pong"), head_dim 64.
// Same arithmetic as attention_tc.cu (single-pass online softmax with lazy rescaling; optional causal + ALiBi; NP = 2 runs
// the hi/lo three-product scheme). What changes is the schedule: the per-key-block chain
//     QK^T MMA -> tcgen05.ld -> max/exp -> P to smem -> PV MMA
// is latency-bound for one tile, so a CTA owns a PAIR of adjacent 128-query tiles of the same (sequence, head):
//   * two softmax warpgroups (4 warps each, thread = one query row x 64 keys: no cross-thread max exchange needed),
//   * the MMA warp interleaves the two tiles' QK^T / PV issues, so one tile's tensor work hides the other's softmax,
//   * both tiles consume the same K / V blocks from one TMA ring (half the L2->smem traffic per query row).
// Key blocks are 64 wide so that both tiles' S, O and P fit in the 512 TMEM columns even with hi/lo planes.
// TMEM: per tile two 64-column S buffers, a 64-column O accumulator and P as packed fp16 pairs (32 columns hi + 32 lo): 256 columns per
// tile, all 512 in use. P reaches the PV MMA through TMEM (TS-form tcgen05.mma), so shared memory only holds Q and the K/V ring.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace pg {

namespace {

constexpr int QT = 128, KB = 64;
constexpr uint32_t QTILE = 16384;  // 128 rows x 64 fp16
constexpr uint32_t KTILE = 8192;   // 64 rows x 64 fp16
constexpr int NSLOT = 8;
constexpr uint32_t TMEM_COLS = 512;

template <int NP>
struct Smem2 {
  static constexpr uint32_t Q = 0;                                 // [2 tiles][NP]
  static constexpr uint32_t KV = 2 * NP * QTILE;                   // [NSLOT][NP]
  static constexpr uint32_t BAR = KV + NSLOT * NP * KTILE;
  static constexpr uint32_t TOTAL = BAR + 512 + 1024;
};

struct Attn2Params {
  int B, T, heads, nqt, npairs, nkb;
  int d;
  long long lo_off;
  __half* out; long long ldo; long long out_lo_off;
  int causal;
  const float* alibi_slopes;
};

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t cvt2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

template <int NP>
__global__ void __launch_bounds__(384, 1) attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                          const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  using L = Smem2<NP>;
  uint8_t* sQ = smem + L::Q;
  uint8_t* sKV = smem + L::KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;            // [NSLOT]
  uint64_t* kv_empty = kv_full + NSLOT;    // [NSLOT]
  uint64_t* s_full = kv_empty + NSLOT;     // [tile][buf]
  uint64_t* s_empty = s_full + 4;          // [tile][buf]
  uint64_t* p_full = s_empty + 4;          // [tile]
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.heads * p.npairs;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
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

  auto nkeys = [&](int j) {  // keys of block j rounded up to the MMA granularity
    const int rem = p.T - j * KB;
    const int n = rem < KB ? rem : KB;
    return (n + 15) & ~15;
  };
  // key blocks tile qt needs: all of them, or (causal) those up to the block holding its last row
  auto nblocks = [&](int qt) {
    if (qt >= p.nqt) return 0;
    if (!p.causal) return p.nkb;
    const int last = (qt * QT + QT - 1 < p.T - 1 ? qt * QT + QT - 1 : p.T - 1) / KB + 1;
    return last < p.nkb ? last : p.nkb;
  };

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int pi = item % p.npairs, bh = item / p.npairs;
        const int h = bh % p.heads, b = bh / p.heads;
        const int row0 = b * p.T;
        const int cq = h * 64, ck = p.d + h * 64, cv = 2 * p.d + h * 64;
        const int n = max(nblocks(2 * pi), nblocks(2 * pi + 1));
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 2 * NP * QTILE);
        for (int t = 0; t < 2; ++t)
          for (int pl = 0; pl < NP; ++pl)
            tma_load_2d(sQ + (t * NP + pl) * QTILE, &tmQ, q_full, cq + pl * static_cast<int>(p.lo_off), row0 + (2 * pi + t) * QT);
        auto load_block = [&](int col, int j) {
          mbar_wait(&kv_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[slot], NP * KTILE);
          for (int pl = 0; pl < NP; ++pl)
            tma_load_2d(sKV + (slot * NP + pl) * KTILE, &tmK, &kv_full[slot], col + pl * static_cast<int>(p.lo_off), row0 + j * KB);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        };
        for (int j = 0; j <= n; ++j) {  // consumption order: K0, K1, V0, K2, V1, ..., V(n-1)
          if (j < n) load_block(ck, j);
          if (j >= 1) load_block(cv, j - 1);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_f16(QT, 64, 0, 1);  // A = P (K-major), B = V (MN-major)
      int slot = 0;
      uint32_t phase = 0;
      uint32_t sblk[2] = {0, 0}, pblk[2] = {0, 0}, oitem[2] = {0, 0};
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int pi = item % p.npairs;
        const int nt[2] = {nblocks(2 * pi), nblocks(2 * pi + 1)};
        const int n = max(nt[0], nt[1]);
        mbar_wait(q_full, it & 1);
        for (int j = 0; j <= n; ++j) {
          if (j < n) {  // S_t = Q_t K_j^T for both tiles
            mbar_wait(&kv_full[slot], phase);
            const uint32_t k_addr = smem_u32(sKV + slot * NP * KTILE);
            const uint32_t idesc_s = make_idesc_f16(QT, nkeys(j), 0, 0);
            for (int t = 0; t < 2; ++t) {
              if (j >= nt[t]) continue;
              const uint32_t buf = sblk[t] & 1;
              mbar_wait(&s_empty[t * 2 + buf], ((sblk[t] >> 1) & 1) ^ 1);
              tc_fence_after();
              const uint32_t q_addr = smem_u32(sQ + t * NP * QTILE);
              const uint32_t tmem_s = tmem_base + t * 256 + buf * 64;
              const uint64_t qh = make_desc_sw128(q_addr, 1024), kh = make_desc_sw128(k_addr, 1024);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, qh + 2 * ks, kh + 2 * ks, idesc_s, ks > 0);
              if (NP == 2) {
                const uint64_t ql = make_desc_sw128(q_addr + QTILE, 1024), kl = make_desc_sw128(k_addr + KTILE, 1024);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, ql + 2 * ks, kh + 2 * ks, idesc_s, 1);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, qh + 2 * ks, kl + 2 * ks, idesc_s, 1);
              }
              umma_commit(&s_full[t * 2 + buf]);
              ++sblk[t];
            }
            umma_commit(&kv_empty[slot]);
            if (++slot == NSLOT) { slot = 0; phase ^= 1; }
            if (j == n - 1) umma_commit(q_empty);  // every QK^T of this pair has been issued
          }
          if (j >= 1) {  // O_t += P_t V_{j-1}
            const int jj = j - 1;
            mbar_wait(&kv_full[slot], phase);
            const uint32_t v_addr = smem_u32(sKV + slot * NP * KTILE);
            const int nks = nkeys(jj) >> 4;
            for (int t = 0; t < 2; ++t) {
              if (jj >= nt[t]) continue;
              mbar_wait(&p_full[t], pblk[t] & 1);
              if (jj == 0) mbar_wait(&o_empty[t], (oitem[t] & 1) ^ 1);
              tc_fence_after();
              const uint32_t tmem_o = tmem_base + t * 256 + 128;
              const uint32_t tmem_p = tmem_base + t * 256 + 192;  // P hi: 32 packed columns (64 keys); lo at +32
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t vh = make_desc_sw128(v_addr + ks * 2048, 1024, 1024);
                umma_f16_ts(tmem_o, tmem_p + ks * 8, vh, idesc_o, (jj > 0 || ks > 0) ? 1u : 0u);  // A = P from TMEM
                if (NP == 2) {
                  const uint64_t vl = make_desc_sw128(v_addr + KTILE + ks * 2048, 1024, 1024);
                  umma_f16_ts(tmem_o, tmem_p + 32 + ks * 8, vh, idesc_o, 1);
                  umma_f16_ts(tmem_o, tmem_p + ks * 8, vl, idesc_o, 1);
                }
              }
              umma_commit(&p_empty[t]);
              ++pblk[t];
              if (jj == nt[t] - 1) {
                umma_commit(&o_full[t]);
                ++oitem[t];
              }
            }
            umma_commit(&kv_empty[slot]);
            if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================= softmax + epilogue: warpgroup t owns tile t of the pair
    const int t = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t sblk = 0, pblk = 0, oitem = 0;
    const uint32_t tmem_t = tmem_base + t * 256 + lane_addr;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int pi = item % p.npairs, bh = item / p.npairs;
      const int h = bh % p.heads, b = bh / p.heads;
      const int qt = 2 * pi + t;
      const int n = nblocks(qt);
      if (n == 0) continue;  // dead tile of an odd tail pair: the MMA warp issues nothing for it either
      const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.f;
      const float slope2 = slope * LOG2E;
      const bool plain = !p.causal && slope == 0.f;
      const bool live = qt * QT + wq * 32 < p.T;  // warps whose rows all lie beyond T only keep the barrier protocol going
      const int qidx = qt * QT + row;
      float m_run = -INFINITY, l = 0.f;
      for (int j = 0; j < n; ++j, ++sblk, ++pblk) {
        const uint32_t buf = sblk & 1;
        mbar_wait(&s_full[t * 2 + buf], (sblk >> 1) & 1);
        tc_fence_after();
        const int valid = p.T - j * KB;                                            // real keys in this block (may exceed KB)
        const int vrow = p.causal ? min(valid, qidx + 1 - j * KB) : valid;          // ... visible to this row
        const int ncols = live ? nkeys(j) : 0;                                      // columns the PV MMA will read
        uint32_t r[2][32];
        if (ncols > 0) tmem_ld_32x32b_x32(tmem_t + buf * 64, r[0]);
        if (ncols > 32) tmem_ld_32x32b_x32(tmem_t + buf * 64 + 32, r[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[t * 2 + buf]);
        float scale = 1.f;
        if (live) {
          const float bias0 = slope2 * static_cast<float>(j * KB);
          float mloc = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < ncols) {
              if (plain && valid - c * 32 >= 32) {
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float x = __uint_as_float(r[c][i]) * LOG2E;
                  r[c][i] = __float_as_uint(x);
                  m4[i & 3] = fmaxf(m4[i & 3], x);
                }
                mloc = fmaxf(mloc, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
              } else if (vrow - c * 32 >= 32) {
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float x = fmaf(__uint_as_float(r[c][i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                  r[c][i] = __float_as_uint(x);
                  m4[i & 3] = fmaxf(m4[i & 3], x);
                }
                mloc = fmaxf(mloc, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  float x = fmaf(__uint_as_float(r[c][i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                  x = (c * 32 + i < vrow) ? x : -INFINITY;
                  r[c][i] = __float_as_uint(x);
                  mloc = fmaxf(mloc, x);
                }
              }
            }
          }
          if (mloc > m_run + 8.f) {  // lazy rescale: raise the running max only when it is exceeded by > 2^8
            scale = ex2a(m_run - mloc);
            m_run = mloc;
          }
          float lsum = 0.f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < ncols) {
              float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float e = ex2a(__uint_as_float(r[c][i]) - m_run);
                l4[i & 3] += e;
                r[c][i] = __float_as_uint(e);
              }
              lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            }
          }
          l = fmaf(l, scale, lsum);
        }
        mbar_wait(&p_empty[t], (pblk & 1) ^ 1);  // PV of the previous block done: O stable, P free
        if (j > 0 && live && __any_sync(0xffffffffu, scale != 1.f)) {
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_t + 128 + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
            tmem_st_32x32b_x32(tmem_t + 128 + c * 32, o);
          }
          tmem_st_wait();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c * 32 < ncols) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float x0 = __uint_as_float(r[c][2 * u]), x1 = __uint_as_float(r[c][2 * u + 1]);
              hi[u] = cvt2(x0, x1);
              if (NP == 2) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
                lo[u] = cvt2(x0 - hf.x, x1 - hf.y);
              }
            }
            tmem_st_32x32b_x16(tmem_t + 192 + c * 16, hi);
            if (NP == 2) tmem_st_32x32b_x16(tmem_t + 224 + c * 16, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- epilogue: O / l -> fp16 hi[/lo], one 128-byte row per thread ----
      mbar_wait(&o_full[t], oitem & 1);
      ++oitem;
      tc_fence_after();
      const float rl = 1.f / l;
      __half* orow = p.out + (static_cast<long long>(b) * p.T + qidx) * p.ldo + h * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        if (live) {
          tmem_ld_32x32b_x32(tmem_t + 128 + c * 32, o);
          tmem_ld_wait();
        }
        if (live && qidx < p.T) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float x0 = __uint_as_float(o[c8 * 8 + 2 * u]) * rl, x1 = __uint_as_float(o[c8 * 8 + 2 * u + 1]) * rl;
              hi[u] = cvt2(x0, x1);
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
              lo[u] = cvt2(x0 - hf.x, x1 - hf.y);
            }
            *reinterpret_cast<uint4*>(orow + c * 32 + c8 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (p.out_lo_off > 0) *reinterpret_cast<uint4*>(orow + p.out_lo_off + c * 32 + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[t]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_attention_tc2(const AttnLaunch& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return PG_OK;
  if (a.nseg != 1 && a.nseg != 3) return set_error(PG_ERR_ARG, "attention: nseg must be 1 or 3");
  if (a.ld % 8 || a.lo_off % 8 || a.ldo % 8 || a.out_lo_off % 8 || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return set_error(PG_ERR_ARG, "attention_tc2: pitches must be multiples of 8 elements and out 16-byte aligned");
  Attn2Params p{};
  p.B = a.B; p.T = a.T; p.heads = a.heads; p.d = a.heads * 64;
  p.nqt = (a.T + QT - 1) / QT; p.npairs = (p.nqt + 1) / 2; p.nkb = (a.T + KB - 1) / KB;
  p.lo_off = a.lo_off; p.out = a.out; p.ldo = a.ldo; p.out_lo_off = a.out_lo_off;
  p.causal = a.causal; p.alibi_slopes = a.alibi_slopes;
  const int np = a.nseg == 3 ? 2 : 1;
  const uint64_t width = static_cast<uint64_t>(3) * p.d * np;
  if (np == 2 && a.lo_off != 3ll * p.d) return set_error(PG_ERR_ARG, "attention_tc2: lo planes must follow the hi planes (lo_off == 3*d)");
  CUtensorMap tmQ, tmK;
  int rc = make_tmap_f16_2d(&tmQ, a.qkv, static_cast<uint64_t>(a.B) * a.T, width, a.ld, 128, 64);
  if (rc) return rc;
  rc = make_tmap_f16_2d(&tmK, a.qkv, static_cast<uint64_t>(a.B) * a.T, width, a.ld, 64, 64);
  if (rc) return rc;
  const long long nitems = static_cast<long long>(a.B) * a.heads * p.npairs;
  const int grid = nitems < num_sms() ? static_cast<int>(nitems) : num_sms();
  static bool attr_set = false;
  if (!attr_set) {
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2<1>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2<2>::TOTAL));
    attr_set = true;
  }
  if (np == 1) attn_tc2_kernel<1><<<grid, 384, Smem2<1>::TOTAL, s>>>(tmQ, tmK, p);
  else attn_tc2_kernel<2><<<grid, 384, Smem2<2>::TOTAL, s>>>(tmQ, tmK, p);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
