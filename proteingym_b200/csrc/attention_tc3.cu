// K4 — tcgen05/TMEM attention, head_dim 64, with P written IN PLACE over the S tile it was computed from.
// Single-pass online softmax with lazy rescaling, optional causal + ALiBi; NP = 2 runs the hi/lo three-product scheme.
// TMEM plan and synchronisation per key block:
//
//   TMEM columns   [0,128) [128,256) [256,384)  a ring of three 128x128 fp32 S tiles;   [384,448) the 128x64 fp32 O accumulator.
//   P(j) = exp2(S(j) - m) is stored as packed fp16 pairs into the SAME 128 columns S(j) occupied (hi pairs in the first 64 columns,
//   lo pairs in the last 64), so P is triple-buffered for free and the PV MMA (TS form: A from TMEM) reads it there.
//
//   Who may touch ring slot b = n % 3 (n = running key-block number), in order:
//     QK(n)  writes S      <- MMA thread, after it issued PV(n-3): tcgen05.mma instructions of one thread execute in issue order,
//                             so no barrier is needed between PV(n-3) reading P(n-3) and QK(n) overwriting it
//     softmax reads S(n)   <- after s_full[b]
//     softmax writes P(n)  <- each thread first loaded its own 64 S columns; the columns it overwrites that belong to the row's
//                             other thread are touched only after the pair barrier both threads pass after their loads
//     PV(n)  reads P(n)    <- after p_full[b]
//   Compared with round 1's first tcgen05 kernel (P in its own TMEM columns) there is no wait for "previous PV finished, P buffer free" (the largest per-block stall in its
//   ncu profile, profiles/ncu_r01_attention_ptmem_summary.txt) and no s_empty hand-back; the softmax warps wait for a finished PV
//   only on the rare blocks that rescale O (pv_done), and in the epilogue (o_full).
//   MMA issue order per work item:  QK0 QK1 QK2 | PV0 QK3 | PV1 QK4 | ...   TMA load order: K0 K1 K2 | V0 K3 | V1 K4 | ...
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace pg {

namespace {

constexpr int QT = 128, KT = 128;
constexpr uint32_t TILE = 16384;  // 128 rows x 64 fp16
constexpr int NSLOT = 4;          // K/V smem ring
constexpr int NBUF = 3;           // S/P TMEM ring
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t O_COL = NBUF * KT;  // 384

template <int NP>
struct Smem3 {
  static constexpr uint32_t Q = 0;
  static constexpr uint32_t KV = NP * TILE;
  static constexpr uint32_t BAR = KV + NSLOT * NP * TILE;
  static constexpr uint32_t TOTAL = BAR + 256 + 2048 + 1024;  // barriers, red[] (2 KiB), alignment slack
};

struct Attn3Params {
  int B, T, heads, nqt, nkb;   // T = full sequence length (prefix + own rows); nqt = query tiles of the own rows; nkb = key blocks of T
  int Tq, s_blocks;            // own rows per sequence; shared-prefix length / 128 (0 = no prefix: Tq == T)
  int d;
  long long lo_off;
  __half* out; long long ldo; long long out_lo_off;
  int causal;
  const float* alibi_slopes;
  int out_fmt; float out_scale;  // common.h operand formats: 1 = fp16 lo plane, 2 = e4m3 [lo8 | hi8] planes for the out_proj GEMM
};

__device__ __forceinline__ float ex2a3(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t cvt2h(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

template <int NP>
__global__ void __launch_bounds__(384, 1) attn_tc3_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmP,
                                                          const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  using L = Smem3<NP>;
  uint8_t* sQ = smem + L::Q;
  uint8_t* sKV = smem + L::KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;              // [NSLOT]
  uint64_t* kv_empty = kv_full + NSLOT;      // [NSLOT]
  uint64_t* s_full = kv_empty + NSLOT;       // [NBUF]  MMA commit: S(n) complete
  uint64_t* p_full = s_full + NBUF;          // [NBUF]  8 softmax warps: P(n) stored
  uint64_t* pv_done = p_full + NBUF;         // [NBUF]  MMA commit: PV(n) complete (O stable up to block n)
  uint64_t* o_full = pv_done + NBUF;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 1);
  __half* redh = reinterpret_cast<__half*>(smem + L::BAR + 256);        // [2 parities][2 halves][128 rows] block maxima (1 KiB)
  float* red = reinterpret_cast<float*>(smem + L::BAR + 256 + 1024);    // [2 halves][128 rows] row sums (1 KiB), separate on purpose

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
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
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
  auto item_nkb = [&](int qt) { return p.causal ? (qt + 1 < p.nkb ? qt + 1 : p.nkb) : p.nkb; };  // QT == KT: diagonal block = qt

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int qt_l = item % p.nqt, bh = item / p.nqt;
        const int qt = qt_l + p.s_blocks;  // tile index within the full sequence
        const int h = bh % p.heads, b = bh / p.heads;
        const int row0 = b * p.Tq;
        const int cq = h * 64, ck = p.d + h * 64, cv = 2 * p.d + h * 64;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, NP * TILE);
        for (int pl = 0; pl < NP; ++pl) tma_load_2d(sQ + pl * TILE, &tm, q_full, cq + pl * static_cast<int>(p.lo_off), row0 + qt_l * QT);
        auto load_block = [&](int col, int j) {  // key block j of the full sequence: shared prefix rows or this sequence's own rows
          mbar_wait(&kv_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&kv_full[slot], NP * TILE);
          const CUtensorMap* src = j < p.s_blocks ? &tmP : &tm;
          const int r = j < p.s_blocks ? j * KT : row0 + (j - p.s_blocks) * KT;
          for (int pl = 0; pl < NP; ++pl)
            tma_load_2d(sKV + (slot * NP + pl) * TILE, src, &kv_full[slot], col + pl * static_cast<int>(p.lo_off), r);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        };
        const int nkb = item_nkb(qt);
        for (int j = 0; j < NBUF && j < nkb; ++j) load_block(ck, j);
        for (int j = 0; j < nkb; ++j) {
          load_block(cv, j);
          if (j + NBUF < nkb) load_block(ck, j + NBUF);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_f16(QT, 64, 0, 1);  // A = P (TMEM), B = V (MN-major)
      int slot = 0;
      uint32_t phase = 0;
      uint32_t qkn = 0, pvn = 0;  // running key-block numbers of the next QK / PV to issue (equal at item boundaries)
      int it = 0;
      const uint32_t q_addr = smem_u32(sQ);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        const int nkb = item_nkb(item % p.nqt + p.s_blocks);
        int jq = 0;  // next key block of this item whose QK has not been issued
        auto issue_qk = [&]() {
          const uint32_t buf = qkn % NBUF;
          mbar_wait(&kv_full[slot], phase);
          tc_fence_after();
          const uint32_t idesc_s = make_idesc_f16(QT, nkeys(jq), 0, 0);
          const uint32_t k_addr = smem_u32(sKV + slot * NP * TILE);
          const uint32_t tmem_s = tmem_base + buf * KT;
          const uint64_t qh = make_desc_sw128(q_addr, 1024), kh = make_desc_sw128(k_addr, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, qh + 2 * ks, kh + 2 * ks, idesc_s, ks > 0);
          if (NP == 2) {
            const uint64_t ql = make_desc_sw128(q_addr + TILE, 1024), kl = make_desc_sw128(k_addr + TILE, 1024);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, ql + 2 * ks, kh + 2 * ks, idesc_s, 1);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_s, qh + 2 * ks, kl + 2 * ks, idesc_s, 1);
          }
          umma_commit(&kv_empty[slot]);
          umma_commit(&s_full[buf]);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          ++qkn;
          ++jq;
          if (jq == nkb) umma_commit(q_empty);  // every QK of this item issued: the Q tile is free once they complete
        };
        while (jq < NBUF && jq < nkb) issue_qk();
        for (int j = 0; j < nkb; ++j) {
          const uint32_t buf = pvn % NBUF;
          mbar_wait(&kv_full[slot], phase);                    // V(j)
          mbar_wait(&p_full[buf], (pvn / NBUF) & 1);            // P(j) stored by all 8 softmax warps
          if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);          // previous item's O has been read out
          tc_fence_after();
          const uint32_t v_addr = smem_u32(sKV + slot * NP * TILE);
          const uint32_t tmem_o = tmem_base + O_COL;
          const uint32_t tmem_p = tmem_base + buf * KT;          // hi pairs: columns [0,64) of the slot; lo pairs: [64,128)
          const int nks = nkeys(j) >> 4;
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t vh = make_desc_sw128(v_addr + ks * 2048, 1024, 1024);
            umma_f16_ts(tmem_o, tmem_p + ks * 8, vh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
            if (NP == 2) {
              const uint64_t vl = make_desc_sw128(v_addr + TILE + ks * 2048, 1024, 1024);
              umma_f16_ts(tmem_o, tmem_p + 64 + ks * 8, vh, idesc_o, 1);
              umma_f16_ts(tmem_o, tmem_p + ks * 8, vl, idesc_o, 1);
            }
          }
          umma_commit(&kv_empty[slot]);
          umma_commit(&pv_done[buf]);
          if (j == nkb - 1) umma_commit(o_full);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          ++pvn;
          if (jq < nkb) issue_qk();                             // QK(j + 3) into the slot PV(j) has just been issued from
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================= softmax + epilogue
    // 8 warps: thread = (query row, 64-key column half g); warps w and w + 4 hold the two halves of the same 32 rows.
    const int wq = warp & 3;
    const int g = (warp - 4) >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t n = 0;  // running key-block number (same count as the MMA thread's)
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int qt = item % p.nqt + p.s_blocks, bh = item / p.nqt;
      const int h = bh % p.heads, b = bh / p.heads;
      const int nkb = item_nkb(qt);
      const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.f;
      const bool plain = !p.causal && slope == 0.f;
      const bool live = qt * QT + wq * 32 < p.T;  // warps whose rows all lie beyond T only keep the barrier protocol going
      const float slope2 = slope * LOG2E;
      float l = 0.f;
      float m_run = -INFINITY;  // log2 domain
      for (int j = 0; j < nkb; ++j, ++n) {
        const uint32_t buf = n % NBUF;
        const uint32_t tmem_s = tmem_base + lane_addr + buf * KT;
        mbar_wait(&s_full[buf], (n / NBUF) & 1);
        tc_fence_after();
        const int valid = p.T - j * KT - g * 64;
        const int vrow = (p.causal && j == qt) ? min(valid, row + 1 - g * 64) : valid;
        const float bias0 = slope2 * static_cast<float>(j * KT + g * 64);
        const int ncols = live ? nkeys(j) - g * 64 : 0;  // columns of this thread's half the PV MMA will read
        uint32_t r[2][32];
        if (ncols > 0) tmem_ld_32x32b_x32(tmem_s + g * 64, r[0]);
        if (ncols > 32) tmem_ld_32x32b_x32(tmem_s + g * 64 + 32, r[1]);
        tmem_ld_wait();
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
          // (2) agree on the block maximum with the thread holding the other 64 columns of this row (same lane, warp +-4).
          //     Passing this barrier also means the partner's tcgen05.ld of its S columns has completed (it waited for the loads
          //     before computing its maximum), so from here on either thread may overwrite any column of the slot with P.
          const __half mh = __float2half_rn(mloc);
          redh[(j & 1) * 256 + g * 128 + row] = mh;
          asm volatile("bar.sync %0, 64;" ::"r"(2 + wq) : "memory");
          const float mblk = fmaxf(__half2float(mh), __half2float(redh[(j & 1) * 256 + (g ^ 1) * 128 + row]));
          // (3) lazy rescale decision (identical in both threads of the row)
          if (mblk > m_run + 8.f) {
            scale = ex2a3(m_run - mblk);  // 0 for the first block (m_run = -inf)
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
                const float e = ex2a3(__uint_as_float(r[c][i]) - m_run);
                l4[i & 3] += e;
                r[c][i] = __float_as_uint(e);
              }
              lsum += (l4[0] + l4[1]) + (l4[2] + l4[3]);
            }
          }
          l = fmaf(l, scale, lsum);
          // (5) rare: the running maximum moved -> bring the O accumulated so far to the new reference. PV(n-1) (and with it every
          //     earlier PV) must have completed; PV(n) cannot start before this warp arrives on p_full below.
          if (j > 0 && __any_sync(0xffffffffu, scale != 1.f)) {
            mbar_wait(&pv_done[(n - 1) % NBUF], ((n - 1) / NBUF) & 1);
            tc_fence_after();
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + g * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
            tmem_st_32x32b_x32(tmem_base + lane_addr + O_COL + g * 32, o);
          }
          // (6) P -> fp16 pairs, in place over the S slot
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 32 < ncols) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const float x0 = __uint_as_float(r[c][2 * u]), x1 = __uint_as_float(r[c][2 * u + 1]);
                hi[u] = cvt2h(x0, x1);
                if (NP == 2) {
                  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
                  lo[u] = cvt2h(x0 - hf.x, x1 - hf.y);
                }
              }
              tmem_st_32x32b_x16(tmem_s + g * 32 + c * 16, hi);
              if (NP == 2) tmem_st_32x32b_x16(tmem_s + 64 + g * 32 + c * 16, lo);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      }
      // ---- epilogue: O / l -> fp16 hi[/lo]; this thread owns 32 of the 64 head-dim columns ----
      red[g * 128 + row] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // row sums of both halves visible
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
        const long long orow_idx = static_cast<long long>(b) * p.Tq + (qidx - p.s_blocks * QT);
        __half* orow = p.out + orow_idx * p.ldo + h * 64 + g * 32;
        uint32_t hi[16], lo[16];
        float lf[32];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float x0 = __uint_as_float(o[2 * u]) * rl, x1 = __uint_as_float(o[2 * u + 1]) * rl;
          hi[u] = cvt2h(x0, x1);
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
          lf[2 * u] = x0 - hf.x;
          lf[2 * u + 1] = x1 - hf.y;
          lo[u] = cvt2h(lf[2 * u], lf[2 * u + 1]);
        }
        uint4* d4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int u = 0; u < 4; ++u) d4[u] = make_uint4(hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        if (p.out_fmt == 1) {
          uint4* l4 = reinterpret_cast<uint4*>(orow + p.out_lo_off);
#pragma unroll
          for (int u = 0; u < 4; ++u) l4[u] = make_uint4(lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
        } else if (p.out_fmt == 2) {
          uint8_t* f8 = reinterpret_cast<uint8_t*>(p.out + orow_idx * p.ldo + p.out_lo_off) + h * 64 + g * 32;
          const float sh = p.out_scale, sl = p.out_scale * 2048.f;
          uint32_t w8l[8], w8h[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi[2 * u]));
            const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi[2 * u + 1]));
            w8h[u] = pack4_e4m3(h01.x * sh, h01.y * sh, h23.x * sh, h23.y * sh);
            w8l[u] = pack4_e4m3(lf[4 * u] * sl, lf[4 * u + 1] * sl, lf[4 * u + 2] * sl, lf[4 * u + 3] * sl);
          }
          uint4* l4 = reinterpret_cast<uint4*>(f8);
          uint4* h4 = reinterpret_cast<uint4*>(f8 + p.d);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            l4[u] = make_uint4(w8l[4 * u], w8l[4 * u + 1], w8l[4 * u + 2], w8l[4 * u + 3]);
            h4[u] = make_uint4(w8h[4 * u], w8h[4 * u + 1], w8h[4 * u + 2], w8h[4 * u + 3]);
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // red[] / redh[] are rewritten by the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_attention_tc3(const AttnLaunch& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return PG_OK;
  if (a.perm_C) return set_error(PG_ERR_UNSUPPORTED, "attention: the column-attention row order is only in the tcgen05 kernel (attention_tc4.cu)");
  if (a.nseg != 1 && a.nseg != 3) return set_error(PG_ERR_ARG, "attention: nseg must be 1 or 3");
  if (a.ld % 8 || a.lo_off % 8 || a.ldo % 8 || a.out_lo_off % 8 || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return set_error(PG_ERR_ARG, "attention_tc3: pitches must be multiples of 8 elements and out 16-byte aligned");
  if (a.q_begin != 0) return set_error(PG_ERR_ARG, "attention_tc3: q_begin is only supported by the mma.sync kernel");
  if (a.prefix && (!a.causal || a.prefix_len <= 0 || a.prefix_len % KT))
    return set_error(PG_ERR_ARG, "attention_tc3: a shared prefix needs causal attention and a length that is a multiple of 128");
  Attn3Params p{};
  p.B = a.B; p.heads = a.heads; p.d = a.heads * 64;
  p.Tq = a.T; p.s_blocks = a.prefix ? a.prefix_len / KT : 0;
  p.T = a.T + p.s_blocks * KT;
  p.nqt = (a.T + QT - 1) / QT; p.nkb = (p.T + KT - 1) / KT;
  p.lo_off = a.lo_off; p.out = a.out; p.ldo = a.ldo; p.out_lo_off = a.out_lo_off;
  p.causal = a.causal; p.alibi_slopes = a.alibi_slopes;
  p.out_fmt = a.out_fmt < 0 ? (a.out_lo_off > 0 ? 1 : 0) : a.out_fmt;
  p.out_scale = a.out_scale;
  if (p.out_fmt > 2 || (p.out_fmt >= 1 && a.out_lo_off <= 0) || (p.out_fmt == 2 && !(a.out_scale > 0.f)))
    return set_error(PG_ERR_ARG, "attention_tc3: bad output format");
  const int np = a.nseg == 3 ? 2 : 1;
  const uint64_t width = static_cast<uint64_t>(3) * p.d * np;
  if (np == 2 && a.lo_off != 3ll * p.d) return set_error(PG_ERR_ARG, "attention_tc3: lo planes must follow the hi planes (lo_off == 3*d)");
  CUtensorMap tm, tmP;
  int rc = make_tmap_f16_2d(&tm, a.qkv, static_cast<uint64_t>(a.B) * a.T, width, a.ld, 128, 64);
  if (rc) return rc;
  tmP = tm;
  if (a.prefix) {
    rc = make_tmap_f16_2d(&tmP, a.prefix, static_cast<uint64_t>(a.prefix_len), width, a.ld, 128, 64);
    if (rc) return rc;
  }
  const long long nitems = static_cast<long long>(a.B) * a.heads * p.nqt;
  const int grid = nitems < num_sms() ? static_cast<int>(nitems) : num_sms();
  int dev = 0;
  PG_CUDA_OK(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (dev < 64 && !attr_set[dev]) {
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem3<1>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem3<2>::TOTAL));
    attr_set[dev] = true;
  }
  if (np == 1) attn_tc3_kernel<1><<<grid, 384, Smem3<1>::TOTAL, s>>>(tm, tmP, p);
  else attn_tc3_kernel<2><<<grid, 384, Smem3<2>::TOTAL, s>>>(tm, tmP, p);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
