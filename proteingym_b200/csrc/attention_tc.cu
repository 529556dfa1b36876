// K4 — tcgen05/TMEM fused multi-head self-attention (bidirectional, or causal + ALiBi for Tranception), head_dim 64.
// Reference arithmetic: esm/multihead_attention.py:357 (QK^T), :379 (fp32 softmax), :387 (PV); q pre-scaled / pre-rotated.
// This file's kernel keeps P in its OWN TMEM columns (or, PG_ATTN_P_SMEM=1, in swizzled shared memory); the model's default since
// the end of round 1 is attention_tc3.cu (P in place over S). launch_attention_tc() below is the dispatch both go through.
//
// One persistent CTA per SM (384 threads); work item = (sequence, head, 128-query tile):
//   warp 0     TMA producer: Q tile (128x64) and a ring of 128-key K / V tiles (SWIZZLE_128B), in MMA consumption order
//   warp 1     MMA issuer (one thread):  S = Q K^T  (tcgen05.mma M128 N<=128 K16, K-major B)  -> TMEM S[2] (fp32)
//                                         O += P V   (A = P from TMEM, TS form; B = V MN-major)  -> TMEM O (128x64 fp32)
//   warp 2     TMEM allocator
//   warps 4-11 softmax: thread = (query row, 64-key half) (tcgen05.ld 32x32b), exp2 in fp32, P -> fp16 (hi[, lo]) -> tcgen05.st
//
// Softmax (template ONLINE): single pass with lazy rescaling — the running row maximum is raised only when a block exceeds it by
// more than 8 (log2 units), so O in TMEM is corrected only on those rare blocks. ONLINE = false (PG_ATTN_TWO_PASS=1) is the exact
// two-pass variant kept for cross-checks: pass A runs QK^T (hi*hi only) for the row maxima, pass B recomputes S and accumulates.
// NP == 2 (f16x3 parity mode): Q,K,V arrive as hi|lo planes, S = QhKh + QlKh + QhKl, O = PhVh + PlVh + PhVl.
// Roofline: tensor/MUFU bound; algorithmic FLOPs = 4*T^2*64 per (sequence, head).
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace pg {

namespace {

constexpr int QT = 128, KT = 128;
constexpr uint32_t TILE = 16384;  // 128 rows x 64 fp16
constexpr int NSLOT = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t S_COL0 = 0, O_COL = 256, P_COL = 320, PLO_COL = 384;  // TMEM columns (P: fp16 pairs, 64 columns per plane)

template <int NP>
struct Smem {
  static constexpr uint32_t Q = 0;
  static constexpr uint32_t KV = NP * TILE;
  static constexpr uint32_t P = KV + NSLOT * NP * TILE;
  static constexpr uint32_t BAR = P + NP * 2 * TILE;
  static constexpr uint32_t TOTAL = BAR + 256 + 1024 + 1024;
};

struct AttnTcParams {
  int B, T, heads, nqt, nkb;
  int d;                  // heads * 64
  long long lo_off;       // column offset of the lo planes in qkv
  __half* out; long long ldo; long long out_lo_off;
  int causal;                 // keys > query masked (Tranception, model_pytorch.py:162-165); key blocks beyond the diagonal skipped
  const float* alibi_slopes;  // [heads] or null: score += slope_h * key_index (model_pytorch.py:167-168, :373-380)
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {  // 2^x, one MUFU; inputs here are <= ~0, tiny results flush to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float lo_elem, float hi_elem) {  // packed RN convert: {hi_elem, lo_elem}
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

template <int NP, bool ONLINE, bool PTMEM>
__global__ void __launch_bounds__(384, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  using L = Smem<NP>;
  uint8_t* sQ = smem + L::Q;
  uint8_t* sKV = smem + L::KV;
  uint8_t* sP = smem + L::P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;            // [NSLOT]
  uint64_t* kv_empty = kv_full + NSLOT;    // [NSLOT]
  uint64_t* s_full = kv_empty + NSLOT;     // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 1);
  float* red = reinterpret_cast<float*>(smem + L::BAR + 256);  // [2][128] cross-warpgroup row reductions

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.heads * p.nqt;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
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
    const int rem = p.T - j * KT;
    const int n = rem < KT ? rem : KT;
    return (n + 15) & ~15;
  };

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int qt = item % p.nqt, bh = item / p.nqt;
        const int h = bh % p.heads, b = bh / p.heads;
        const int row0 = b * p.T;
        const int cq = h * 64, ck = p.d + h * 64, cv = 2 * p.d + h * 64;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, NP * TILE);
        for (int pl = 0; pl < NP; ++pl) tma_load_2d(sQ + pl * TILE, &tm, q_full, cq + pl * static_cast<int>(p.lo_off), row0 + qt * QT);
        auto load_block = [&](int col, int j, int planes) {
          mbar_wait(&kv_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[slot], planes * TILE);
          for (int pl = 0; pl < planes; ++pl)
            tma_load_2d(sKV + (slot * NP + pl) * TILE, &tm, &kv_full[slot], col + pl * static_cast<int>(p.lo_off), row0 + j * KT);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        };
        const int nkb = p.causal ? (qt + 1 < p.nkb ? qt + 1 : p.nkb) : p.nkb;  // QT == KT: the diagonal block is block qt
        if (!ONLINE)
          for (int j = 0; j < nkb; ++j) load_block(ck, j, 1);  // pass A: K hi only
        load_block(ck, 0, NP);                                // pass B: K0, then (K_{j+1}, V_j) ...
        for (int j = 0; j < nkb; ++j) {
          if (j + 1 < nkb) load_block(ck, j + 1, NP);
          load_block(cv, j, NP);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_f16(QT, 64, 0, 1);  // A = P (K-major), B = V (MN-major)
      int slot = 0;
      uint32_t phase = 0;
      uint32_t sblk = 0, pblk = 0;
      int it = 0;
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        const int qt_m = item % p.nqt;
        const int nkb = p.causal ? (qt_m + 1 < p.nkb ? qt_m + 1 : p.nkb) : p.nkb;
        auto issue_qk = [&](int j, bool full) {
          const uint32_t buf = sblk & 1;
          mbar_wait(&kv_full[slot], phase);
          mbar_wait(&s_empty[buf], ((sblk >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t idesc_s = make_idesc_f16(QT, nkeys(j), 0, 0);
          const uint32_t k_addr = smem_u32(sKV + slot * NP * TILE);
          const uint32_t tmem_s = tmem_base + S_COL0 + buf * KT;
          const uint64_t qh = make_desc_sw128(q_addr, 1024), kh = make_desc_sw128(k_addr, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, qh + 2 * ks, kh + 2 * ks, idesc_s, ks > 0);
          if (NP == 2 && full) {
            const uint64_t ql = make_desc_sw128(q_addr + TILE, 1024), kl = make_desc_sw128(k_addr + TILE, 1024);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, ql + 2 * ks, kh + 2 * ks, idesc_s, 1);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, qh + 2 * ks, kl + 2 * ks, idesc_s, 1);
          }
          umma_commit(&kv_empty[slot]);
          umma_commit(&s_full[buf]);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          ++sblk;
        };
        if (!ONLINE)
          for (int j = 0; j < nkb; ++j) issue_qk(j, false);  // pass A (row max only)
        issue_qk(0, true);
        for (int j = 0; j < nkb; ++j) {
          if (j + 1 < nkb) issue_qk(j + 1, true);
          else umma_commit(q_empty);  // all QK MMAs of this item issued: Q tile free once they complete
          mbar_wait(&kv_full[slot], phase);
          mbar_wait(p_full, pblk & 1);
          if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(sKV + slot * NP * TILE);
          const uint32_t tmem_o = tmem_base + O_COL;
          const int nks = nkeys(j) >> 4;
          for (int ks = 0; ks < nks; ++ks) {
            // A: P[128 x 16 keys] inside 64-key swizzle atoms of 16 KiB; B: V[16 keys x 64] = 2 KiB further per step
            const uint64_t vh = make_desc_sw128(v_addr + ks * 2048, 1024, 1024);
            if (PTMEM) {
              // A = P straight from TMEM (128 lanes x 16 fp16 = 8 packed columns per k-step): no smem round trip, no proxy fence
              umma_f16_ts(tmem_o, tmem_base + P_COL + ks * 8, vh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
              if (NP == 2) {
                const uint64_t vl = make_desc_sw128(v_addr + TILE + ks * 2048, 1024, 1024);
                umma_f16_ts(tmem_o, tmem_base + PLO_COL + ks * 8, vh, idesc_o, 1);
                umma_f16_ts(tmem_o, tmem_base + P_COL + ks * 8, vl, idesc_o, 1);
              }
            } else {
              const uint32_t pa = p_addr + (ks >> 2) * TILE + (ks & 3) * 32;
              const uint64_t ph = make_desc_sw128(pa, 1024);
              umma_f16(tmem_o, ph, vh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
              if (NP == 2) {
                const uint64_t pl = make_desc_sw128(pa + 2 * TILE, 1024);
                const uint64_t vl = make_desc_sw128(v_addr + TILE + ks * 2048, 1024, 1024);
                umma_f16(tmem_o, pl, vh, idesc_o, 1);
                umma_f16(tmem_o, ph, vl, idesc_o, 1);
              }
            }
          }
          umma_commit(&kv_empty[slot]);
          umma_commit(p_empty);
          if (j == nkb - 1) umma_commit(o_full);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          ++pblk;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================= softmax + epilogue
    // 8 warps: thread = (query row, 64-key column half g). Two warps per scheduler hide each other's latencies.
    const int wq = warp & 3;
    const int g = (warp - 4) >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t sblk = 0, pblk = 0;
    int it = 0;
    uint8_t* prow = sP + g * TILE + row * 128;  // this thread's row inside its 64-key swizzle atom
    const int sw = row & 7;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int qt = item % p.nqt, bh = item / p.nqt;
      const int h = bh % p.heads, b = bh / p.heads;
      const int nkb = p.causal ? (qt + 1 < p.nkb ? qt + 1 : p.nkb) : p.nkb;
      const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.f;
      const bool plain = !p.causal && slope == 0.f;  // ESM path: no per-row limit, no bias
      // A warp whose 32 query rows all lie beyond T (the tail tile of T = 514 keeps 2 rows of 128) only keeps the barrier
      // protocol going: no TMEM loads, no exponentials, no P writes (its P rows feed O rows that are never stored).
      const bool live = qt * QT + wq * 32 < p.T;
      float l = 0.f;
      if (ONLINE) {
        // ---- single pass, online softmax with lazy rescaling: the running maximum is only raised when a block exceeds it by
        // more than 8 (log2 units), so P <= 2^8 and O / l in TMEM need a correction step only on those (rare) blocks ----
        const float slope2 = slope * LOG2E;
        float m_run = -INFINITY;  // log2 domain
        __half* redh = reinterpret_cast<__half*>(red);  // [2 parities][2 halves][128 rows]; fp16 is enough for a stabiliser
        for (int j = 0; j < nkb; ++j, ++sblk, ++pblk) {
          const uint32_t buf = sblk & 1;
          mbar_wait(&s_full[buf], (sblk >> 1) & 1);
          tc_fence_after();
          const int valid = p.T - j * KT - g * 64;
          const int vrow = (p.causal && j == qt) ? min(valid, row + 1 - g * 64) : valid;
          const float bias0 = slope2 * static_cast<float>(j * KT + g * 64);
          const int ncols = live ? nkeys(j) - g * 64 : 0;
          uint32_t r[2][32];
          if (ncols > 0) tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL0 + buf * KT + g * 64, r[0]);
          if (ncols > 32) tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL0 + buf * KT + g * 64 + 32, r[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[buf]);
          float scale = 1.f;
          if (live) {
            // (1) scores to the log2 domain (+ ALiBi, masks) and this thread's block maximum
            float mloc = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c * 32 < ncols) {
                if (plain && valid - c * 32 >= 32) {
                  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const float t = __uint_as_float(r[c][i]) * LOG2E;
                    r[c][i] = __float_as_uint(t);
                    m4[i & 3] = fmaxf(m4[i & 3], t);
                  }
                  mloc = fmaxf(mloc, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
                } else if (vrow - c * 32 >= 32) {  // ALiBi, every column of this chunk visible to this row: no masks
                  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const float t = fmaf(__uint_as_float(r[c][i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                    r[c][i] = __float_as_uint(t);
                    m4[i & 3] = fmaxf(m4[i & 3], t);
                  }
                  mloc = fmaxf(mloc, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    float t = fmaf(__uint_as_float(r[c][i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                    t = (c * 32 + i < vrow) ? t : -INFINITY;
                    r[c][i] = __float_as_uint(t);
                    mloc = fmaxf(mloc, t);
                  }
                }
              }
            }
            // (2) agree on the block maximum with the thread holding the other 64 columns of this row (same lane, warp +-4)
            const __half mh = __float2half_rn(mloc);
            redh[(j & 1) * 256 + g * 128 + row] = mh;
            asm volatile("bar.sync %0, 64;" ::"r"(2 + wq) : "memory");
            const float mblk = fmaxf(__half2float(mh), __half2float(redh[(j & 1) * 256 + (g ^ 1) * 128 + row]));
            // (3) lazy rescale decision (identical in both threads of the row)
            if (mblk > m_run + 8.f) {
              scale = ex2_approx(m_run - mblk);  // 0 for the first block (m_run = -inf)
              m_run = mblk;
            }
            // (4) P = 2^(t - m_run), row-sum
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c * 32 < ncols) {
                float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float e = ex2_approx(__uint_as_float(r[c][i]) - m_run);
                  l4[i & 3] += e;
                  r[c][i] = __float_as_uint(e);
                }
                lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
              }
            }
            l = fmaf(l, scale, lsum);
          }
          mbar_wait(p_empty, (pblk & 1) ^ 1);  // PV of the previous block has completed: O is stable, P is free
          if (j > 0 && live && __any_sync(0xffffffffu, scale != 1.f)) {
            tc_fence_after();
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + g * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
            tmem_st_32x32b_x32(tmem_base + lane_addr + O_COL + g * 32, o);
            tmem_st_wait();
          }
          if (PTMEM) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c * 32 < ncols) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                  const float x0 = __uint_as_float(r[c][2 * u]), x1 = __uint_as_float(r[c][2 * u + 1]);
                  hi[u] = cvt_f16x2(x0, x1);
                  if (NP == 2) {
                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
                    lo[u] = cvt_f16x2(x0 - hf.x, x1 - hf.y);
                  }
                }
                tmem_st_32x32b_x16(tmem_base + lane_addr + P_COL + g * 32 + c * 16, hi);
                if (NP == 2) tmem_st_32x32b_x16(tmem_base + lane_addr + PLO_COL + g * 32 + c * 16, lo);
              }
            }
            tmem_st_wait();
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
  #pragma unroll
              for (int q8 = 0; q8 < 4; ++q8) {
                if (c * 32 + q8 * 8 < ncols) {
                  uint32_t hi[4], lo[4];
  #pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float x0 = __uint_as_float(r[c][q8 * 8 + 2 * u]), x1 = __uint_as_float(r[c][q8 * 8 + 2 * u + 1]);
                    hi[u] = cvt_f16x2(x0, x1);
                    if (NP == 2) {
                      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
                      lo[u] = cvt_f16x2(x0 - hf.x, x1 - hf.y);
                    }
                  }
                  uint8_t* dst = prow + (((c * 4 + q8) ^ sw) << 4);
                  *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  if (NP == 2) *reinterpret_cast<uint4*>(dst + 2 * TILE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
              }
            }
          fence_async_smem();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
        }
      } else {
      // ---- pass A: row max over this thread's columns ----
      float m = -INFINITY;
      for (int j = 0; j < nkb; ++j, ++sblk) {
        const uint32_t buf = sblk & 1;
        mbar_wait(&s_full[buf], (sblk >> 1) & 1);
        tc_fence_after();
        const int valid = p.T - j * KT - g * 64;  // valid keys in this thread's 64 columns (may be <= 0 or >= 64)
        // causal: this row may look at keys <= its own index, i.e. at most vrow columns of this thread's half
        const int vrow = (p.causal && j == qt) ? min(valid, row + 1 - g * 64) : valid;
        const float kb0 = static_cast<float>(j * KT + g * 64);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          if (c * 32 >= valid || !live) break;
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL0 + buf * KT + g * 64 + c * 32, r);
          tmem_ld_wait();
          if (plain && valid - c * 32 >= 32) {
            float m4[4] = {m, -INFINITY, -INFINITY, -INFINITY};  // four independent chains instead of one 32-deep FMNMX chain
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[i]));
            m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < vrow) m = fmaxf(m, fmaf(slope, kb0 + static_cast<float>(c * 32 + i), __uint_as_float(r[i])));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
      }
      red[g * 128 + row] = m;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      m = fmaxf(m, red[(g ^ 1) * 128 + row]);
      const float m2 = m * LOG2E;
      // ---- pass B: P = exp(S - max) ----
      const float slope2 = slope * LOG2E;
      for (int j = 0; j < nkb; ++j, ++sblk, ++pblk) {
        const uint32_t buf = sblk & 1;
        mbar_wait(&s_full[buf], (sblk >> 1) & 1);
        tc_fence_after();
        const int valid = p.T - j * KT - g * 64;
        const int vrow = (p.causal && j == qt) ? min(valid, row + 1 - g * 64) : valid;
        const float bias0 = fmaf(slope2, static_cast<float>(j * KT + g * 64), -m2);  // slope*key_index*log2e - max*log2e
        const int ncols = live ? nkeys(j) - g * 64 : 0;  // columns the PV MMA will read from this thread's half
        uint32_t r[2][32];
        if (ncols > 0) tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL0 + buf * KT + g * 64, r[0]);
        if (ncols > 32) tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL0 + buf * KT + g * 64 + 32, r[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[buf]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c * 32 < ncols) {
            if (plain && valid - c * 32 >= 32) {
              float l4[4] = {0.f, 0.f, 0.f, 0.f};  // break the 32-deep FADD dependency chain (fixed association -> deterministic)
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float e = ex2_approx(fmaf(__uint_as_float(r[c][i]), LOG2E, -m2));
                l4[i & 3] += e;
                r[c][i] = __float_as_uint(e);
              }
              l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float arg = fmaf(__uint_as_float(r[c][i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                const float e = (c * 32 + i < vrow) ? ex2_approx(arg) : 0.f;
                l += e;
                r[c][i] = __float_as_uint(e);
              }
            }
          }
        }
        mbar_wait(p_empty, (pblk & 1) ^ 1);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {  // 8 keys = one 16-byte chunk of the 128-byte swizzled row
            if (c * 32 + q8 * 8 < ncols) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float x0 = __uint_as_float(r[c][q8 * 8 + 2 * u]), x1 = __uint_as_float(r[c][q8 * 8 + 2 * u + 1]);
                hi[u] = cvt_f16x2(x0, x1);
                if (NP == 2) {
                  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
                  lo[u] = cvt_f16x2(x0 - hf.x, x1 - hf.y);
                }
              }
              uint8_t* dst = prow + (((c * 4 + q8) ^ sw) << 4);
              *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (NP == 2) *reinterpret_cast<uint4*>(dst + 2 * TILE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      }
      // ---- epilogue: O / l -> fp16 hi[/lo]; this thread owns 32 of the 64 head-dim columns ----
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone has read the pass-A maxima
      red[g * 128 + row] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      l += red[(g ^ 1) * 128 + row];
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      uint32_t o[32];
      if (live) {
        tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + g * 32, o);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      const int qidx = qt * QT + row;
      if (live && qidx < p.T) {
        const float rl = 1.f / l;
        __half* orow = p.out + (static_cast<long long>(b) * p.T + qidx) * p.ldo + h * 64 + g * 32;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float x0 = __uint_as_float(o[2 * u]) * rl, x1 = __uint_as_float(o[2 * u + 1]) * rl;
          hi[u] = cvt_f16x2(x0, x1);
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
          lo[u] = cvt_f16x2(x0 - hf.x, x1 - hf.y);
        }
        uint4* d4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int u = 0; u < 4; ++u) d4[u] = make_uint4(hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        if (p.out_lo_off > 0) {
          uint4* l4 = reinterpret_cast<uint4*>(orow + p.out_lo_off);
#pragma unroll
          for (int u = 0; u < 4; ++u) l4[u] = make_uint4(lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // red[] reuse by the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// The model's dispatch (pg_attention impl 0): which tcgen05 kernel runs.
int launch_attention_tc(const AttnLaunch& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return PG_OK;
  // PG_ATTN_PAIR=1 selects the tile-pair ("ping-pong") kernel of attention_tc2.cu (on par in f16, slower in f16x3).
  static const bool pair = getenv("PG_ATTN_PAIR") && getenv("PG_ATTN_PAIR")[0] == '1';
  if (pair) return launch_attention_tc2(a, s);
  // Default: attention_tc3.cu (P stored in place over S, three-slot TMEM ring, no wait for the previous PV): equal in f16, +4.5 % in
  // f16x3 (B=128, T=514, H=20: 0.868 vs 0.907 ms). PG_ATTN_INPLACE=0 keeps this file's kernel (P in its own TMEM columns).
  static const bool inplace = !(getenv("PG_ATTN_INPLACE") && getenv("PG_ATTN_INPLACE")[0] == '0');
  if (inplace && a.q_begin == 0) return launch_attention_tc3(a, s);
  return launch_attention_tc_own(a, s);
}

// This file's kernel (pg_attention impl 2).
int launch_attention_tc_own(const AttnLaunch& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return PG_OK;
  if (a.nseg != 1 && a.nseg != 3) return set_error(PG_ERR_ARG, "attention: nseg must be 1 or 3");
  if (a.ld % 8 || a.lo_off % 8 || a.ldo % 8 || a.out_lo_off % 8 || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return set_error(PG_ERR_ARG, "attention_tc: pitches must be multiples of 8 elements and out 16-byte aligned");
  AttnTcParams p{};
  p.B = a.B; p.T = a.T; p.heads = a.heads; p.d = a.heads * 64;
  p.nqt = (a.T + QT - 1) / QT; p.nkb = (a.T + KT - 1) / KT;
  // A last query tile holding only a few rows (T = 514 -> 2 of 128) would cost a full work item; those rows go to the
  // mma.sync kernel instead, which runs several CTAs per SM next to this persistent kernel's tail.
  // Measured (B200, T = 514): the serialised mma.sync tail launch costs as much as the wasted tile it removes, so the split is
  // off by default; PG_ATTN_SPLIT_TAIL=1 enables it for experiments.
  const int tail = a.T % QT;
  static const bool want_split = getenv("PG_ATTN_SPLIT_TAIL") != nullptr;
  const bool split_tail = want_split && a.T > QT && tail > 0 && tail <= 16;
  if (split_tail) p.nqt -= 1;
  p.lo_off = a.lo_off; p.out = a.out; p.ldo = a.ldo; p.out_lo_off = a.out_lo_off;
  p.causal = a.causal; p.alibi_slopes = a.alibi_slopes;
  const int np = a.nseg == 3 ? 2 : 1;
  const uint64_t width = static_cast<uint64_t>(3) * p.d * np;
  if (np == 2 && a.lo_off != 3ll * p.d) return set_error(PG_ERR_ARG, "attention_tc: lo planes must follow the hi planes (lo_off == 3*d)");
  CUtensorMap tm;
  int rc = make_tmap_f16_2d(&tm, a.qkv, static_cast<uint64_t>(a.B) * a.T, width, a.ld, 128, 64);
  if (rc) return rc;
  const long long nitems = static_cast<long long>(a.B) * a.heads * p.nqt;
  const int grid = nitems < num_sms() ? static_cast<int>(nitems) : num_sms();
  static bool attr_set = false;
  if (!attr_set) {
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<1>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<1>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<1>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::TOTAL));
    attr_set = true;
  }
  // PG_ATTN_TWO_PASS=1 selects the exact two-pass softmax variant (kept for cross-checking); default is the single-pass kernel.
  static const bool two_pass = getenv("PG_ATTN_TWO_PASS") != nullptr;
  // PG_ATTN_P_SMEM=1: hand P to the PV MMA through swizzled shared memory instead of TMEM (the first working variant).
  static const bool p_smem = getenv("PG_ATTN_P_SMEM") != nullptr;
  if (np == 1) {
    if (two_pass) attn_tc_kernel<1, false, false><<<grid, 384, Smem<1>::TOTAL, s>>>(tm, p);
    else if (p_smem) attn_tc_kernel<1, true, false><<<grid, 384, Smem<1>::TOTAL, s>>>(tm, p);
    else attn_tc_kernel<1, true, true><<<grid, 384, Smem<1>::TOTAL, s>>>(tm, p);
  } else {
    if (two_pass) attn_tc_kernel<2, false, false><<<grid, 384, Smem<2>::TOTAL, s>>>(tm, p);
    else if (p_smem) attn_tc_kernel<2, true, false><<<grid, 384, Smem<2>::TOTAL, s>>>(tm, p);
    else attn_tc_kernel<2, true, true><<<grid, 384, Smem<2>::TOTAL, s>>>(tm, p);
  }
  PG_CUDA_OK(cudaGetLastError());
  if (split_tail) {
    AttnLaunch t = a;
    t.q_begin = a.T - tail;
    return launch_attention(t, s);
  }
  return PG_OK;
}

}  // namespace pg
