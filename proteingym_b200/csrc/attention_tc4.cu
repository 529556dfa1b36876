// K4 — tcgen05/TMEM attention, head_dim 64: two CTAs per SM, 128-query tiles x 64-key blocks, one softmax thread per query row.
// Single-pass online softmax with lazy rescaling, optional causal + ALiBi + shared (wild-type) key/value prefix; NP = 2 runs the
// hi/lo three-product scheme. P is written IN PLACE over the S tile it was computed from (as packed fp16 pairs).
//
// Why this shape (profiles/ncu_r01_attention_ptmem_summary.txt, bench r02): the previous kernel (one CTA per SM, 128x128 blocks, the
// 128 columns of a row split over two threads) ran a serial chain per key block — QK^T -> tcgen05.ld -> max -> pair barrier -> exp ->
// tcgen05.st -> PV — at ~4500 cycles per block with the tensor pipe 30 % busy and two warps per scheduler to hide every latency of it.
// Here the chain is shorter (a thread owns the whole 64-key row of its block: row maximum and row sum are thread-local, no barrier,
// no shared-memory exchange) and TWO independent CTAs share the SM (96 KB of shared memory, 256 TMEM columns, <= 168 registers each),
// so one CTA's MMAs fill the tensor pipe while the other CTA's softmax warps run. 192 threads per CTA: warp 0 TMA producer (and
// TMEM allocation), warp 1 MMA issuer, warps 2-5 softmax + epilogue (TMEM lane quadrants 2, 3, 0, 1).
//
//   TMEM columns (256 per CTA)   [0,64) [64,128) [128,192)  a ring of three 128x64 fp32 S tiles;   [192,256) the 128x64 fp32 O tile.
//   P(j) = exp2(S(j) - m): packed fp16 pairs into the SAME 64 columns S(j) occupied (hi pairs in the first 32, lo pairs in the last
//   32), so P is triple-buffered for free and the PV MMA (TS form: A from TMEM) reads it there.
//   Who may touch ring slot b = n % 3 (n = running key-block number), in order:
//     QK(n)  writes S      <- MMA thread, after it issued PV(n-3): tcgen05.mma instructions of one thread execute in issue order,
//                             so no barrier is needed between PV(n-3) reading P(n-3) and QK(n) overwriting it
//     softmax reads S(n)   <- after s_full[b]; writes P(n) over it (each thread only ever touches its own TMEM lane)
//     PV(n)  reads P(n)    <- after p_full[b] (4 softmax warps)
//   The softmax warps wait for a finished PV only on the rare blocks that rescale O (pv_done), and in the epilogue (o_full).
//   MMA issue order per work item:  QK0 QK1 QK2 | PV0 QK3 | PV1 QK4 | ...   TMA load order: K0 K1 K2 | V0 K3 | V1 K4 | ...
//   Work item = (sequence, head, 128-query tile); persistent CTAs, grid = 2 x SMs.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace pg {

namespace {

constexpr int QT = 128, KT = 64;
constexpr uint32_t QTILE = 16384;  // 128 query rows x 64 fp16
constexpr uint32_t TILE = 8192;    // 64 key rows x 64 fp16
constexpr int NSLOT = 4;           // K/V smem ring
constexpr int NBUF = 3;            // S/P TMEM ring
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t O_COL = NBUF * KT;  // 192
constexpr int ATT_THREADS = 192;     // warp 0 TMA (+ TMEM allocation), warp 1 MMA, warps 2-5 softmax; 2 CTAs/SM -> 168 registers each

template <int NP>
struct Smem4 {
  static constexpr uint32_t Q = 0;
  static constexpr uint32_t KV = NP * QTILE;
  static constexpr uint32_t BAR = KV + NSLOT * NP * TILE;
  static constexpr uint32_t TOTAL = BAR + 256 + 1024;  // barriers, alignment slack  (NP = 2: 99.6 KB -> two CTAs per SM)
};

struct Attn4Params {
  int B, T, heads, nqt, nkb;   // T = full sequence length (prefix + own rows); nqt = query tiles of the own rows; nkb = key blocks of T
  int Tq, s_blocks, s_tiles;   // own rows per sequence; shared-prefix length / 64 and / 128 (0 = no prefix: Tq == T)
  int d;
  long long lo_off;
  __half* out; long long ldo; long long out_lo_off;
  int causal;
  const float* alibi_slopes;
  int out_fmt; float out_scale;  // common.h operand formats: 1 = fp16 lo plane, 2 = e4m3 [lo8 | hi8] planes for the out_proj GEMM
  int perm_C;                    // > 0: sequence b is column (b / perm_C, b % perm_C) of an alignment; output rows go to (., r, c) order
  // delta-operand mode (common.h AttnLaunch): out = attention - base_o[row]; full-precision hi / lo copy of row mask_pos[b] to cout[b]
  const __half* base_o; const int* mask_pos; __half* cout; long long ldc; long long c_lo_off;
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

// Softmax arithmetic: packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2), 3-input maxima (FMNMX3), the log2(e) scale folded into the
// exponent's FFMA2, and exponent + row sum + fp16 hi / lo packing of a 32-column chunk in ONE basic block so that the packing of
// earlier pairs issues between the MUFU.EX2 instructions: ~5 instead of ~9 issue slots per score. Same-box A/B against the scalar
// form (profiles/ab_r02_m_*.txt, profiles/ncu_r02_attn_sv3_summary.txt): 268 vs 254 TFLOP/s hi/lo, 443 vs 400 single pass.
template <int NP, int DELTA>
__global__ void __launch_bounds__(ATT_THREADS, 2) attn_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tm,
                                                                  const __grid_constant__ CUtensorMap tmP, const Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  using L = Smem4<NP>;
  uint8_t* sQ = smem + L::Q;
  uint8_t* sKV = smem + L::KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;              // [NSLOT]
  uint64_t* kv_empty = kv_full + NSLOT;      // [NSLOT]
  uint64_t* s_full = kv_empty + NSLOT;       // [NBUF]  MMA commit: S(n) complete
  uint64_t* p_full = s_full + NBUF;          // [NBUF]  4 softmax warps: P(n) stored
  uint64_t* pv_done = p_full + NBUF;         // [NBUF]  MMA commit: PV(n) complete (O stable up to block n)
  uint64_t* o_full = pv_done + NBUF;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = p.B * p.heads * p.nqt;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tm); }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 0) {
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
  // causal: query tile qt (rows 128 qt ..) needs key blocks 0 .. 2 qt + 1 (64 keys each)
  auto item_nkb = [&](int qt) { return p.causal ? (2 * qt + 2 < p.nkb ? 2 * qt + 2 : p.nkb) : p.nkb; };

  // Producer and MMA warps keep their control flow warp-uniform (all 32 lanes wait on the barriers) and predicate only the issue on
  // one elected lane: a loop under `if (lane == 0)` pays an ELECT/branch sequence around every uniform-datapath instruction, and
  // with 24 MMAs per 64-key block in the hi/lo mode the issuing thread, not the tensor pipe, set the pace.
  if (warp == 0) {
    // ================================================================= TMA producer
    int slot = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int qt_l = item % p.nqt, bh = item / p.nqt;
      const int qt = qt_l + p.s_tiles;  // tile index within the full sequence
      const int h = bh % p.heads, b = bh / p.heads;
      const int row0 = b * p.Tq;
      const int cq = h * 64, ck = p.d + h * 64, cv = 2 * p.d + h * 64;
      mbar_wait(q_empty, (it & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, NP * QTILE);
        for (int pl = 0; pl < NP; ++pl) tma_load_2d(sQ + pl * QTILE, &tmQ, q_full, cq + pl * static_cast<int>(p.lo_off), row0 + qt_l * QT);
      }
      __syncwarp();
      auto load_block = [&](int col, int j) {  // key block j of the full sequence: shared prefix rows or this sequence's own rows
        mbar_wait(&kv_empty[slot], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[slot], NP * TILE);
          const CUtensorMap* src = j < p.s_blocks ? &tmP : &tm;
          const int r = j < p.s_blocks ? j * KT : row0 + (j - p.s_blocks) * KT;
          for (int pl = 0; pl < NP; ++pl)
            tma_load_2d(sKV + (slot * NP + pl) * TILE, src, &kv_full[slot], col + pl * static_cast<int>(p.lo_off), r);
        }
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
      };
      const int nkb = item_nkb(qt);
      for (int j = 0; j < NBUF && j < nkb; ++j) load_block(ck, j);
      for (int j = 0; j < nkb; ++j) {
        load_block(cv, j);
        if (j + NBUF < nkb) load_block(ck, j + NBUF);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    constexpr uint32_t idesc_o = make_idesc_f16(QT, 64, 0, 1);  // A = P (TMEM), B = V (MN-major)
    int slot = 0;
    uint32_t phase = 0;
    uint32_t qkn = 0, pvn = 0;  // running key-block numbers of the next QK / PV to issue (equal at item boundaries)
    int it = 0;
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t kv_addr = smem_u32(sKV);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      mbar_wait(q_full, it & 1);
      const int nkb = item_nkb(item % p.nqt + p.s_tiles);
      int jq = 0;  // next key block of this item whose QK has not been issued
      auto issue_qk = [&]() {
        const uint32_t buf = qkn % NBUF;
        mbar_wait(&kv_full[slot], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc_s = make_idesc_f16(QT, nkeys(jq), 0, 0);
          const uint32_t k_addr = kv_addr + slot * NP * TILE;
          const uint32_t tmem_s = tmem_base + buf * KT;
          const uint64_t qh = make_desc_sw128(q_addr, 1024), kh = make_desc_sw128(k_addr, 1024);
          umma_kblock<0>(tmem_s, qh, kh, idesc_s, 0);  // 4 x K16 over the 64 head dims
          if (NP == 2) {
            const uint64_t ql = make_desc_sw128(q_addr + QTILE, 1024), kl = make_desc_sw128(k_addr + TILE, 1024);
            umma_kblock<0>(tmem_s, ql, kh, idesc_s, 1);
            umma_kblock<0>(tmem_s, qh, kl, idesc_s, 1);
          }
          umma_commit(&kv_empty[slot]);
          umma_commit(&s_full[buf]);
          if (jq + 1 == nkb) umma_commit(q_empty);  // every QK of this item issued: the Q tile is free once they complete
        }
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        ++qkn;
        ++jq;
      };
      while (jq < NBUF && jq < nkb) issue_qk();
      for (int j = 0; j < nkb; ++j) {
        const uint32_t buf = pvn % NBUF;
        mbar_wait(&kv_full[slot], phase);                    // V(j)
        mbar_wait(&p_full[buf], (pvn / NBUF) & 1);            // P(j) stored by all 4 softmax warps
        if (j == 0) mbar_wait(o_empty, (it & 1) ^ 1);          // previous item's O has been read out
        tc_fence_after();
        if (elect_one()) {
          const uint32_t v_addr = kv_addr + slot * NP * TILE;
          const uint32_t tmem_o = tmem_base + O_COL;
          const uint32_t tmem_p = tmem_base + buf * KT;          // hi pairs: columns [0,32) of the slot; lo pairs: [32,64)
          const int nks = nkeys(j) >> 4;
          if (nks == 4) {
            umma_pv64<NP, 32>(tmem_o, tmem_p, make_desc_sw128(v_addr, 1024, 1024), make_desc_sw128(v_addr + TILE, 1024, 1024), idesc_o,
                              j > 0 ? 1u : 0u);
          } else {
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t vh = make_desc_sw128(v_addr + ks * 2048, 1024, 1024);
              umma_f16_ts(tmem_o, tmem_p + ks * 8, vh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
              if (NP == 2) {
                const uint64_t vl = make_desc_sw128(v_addr + TILE + ks * 2048, 1024, 1024);
                umma_f16_ts(tmem_o, tmem_p + 32 + ks * 8, vh, idesc_o, 1);
                umma_f16_ts(tmem_o, tmem_p + ks * 8, vl, idesc_o, 1);
              }
            }
          }
          umma_commit(&kv_empty[slot]);
          umma_commit(&pv_done[buf]);
          if (j == nkb - 1) umma_commit(o_full);
        }
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        ++pvn;
        if (jq < nkb) issue_qk();                             // QK(j + 3) into the slot PV(j) has just been issued from
      }
    }
  } else if (warp >= 2) {
    // ================================================================= softmax + epilogue
    // 4 warps: thread = one query row of the tile, all 64 keys of the block (TMEM lane = row), so the row maximum, the row sum and
    // the lazy-rescale decision are thread-local.
    const int wq = warp & 3;  // TMEM lane quadrant this warp may access: warps 2, 3, 4, 5 -> quadrants 2, 3, 0, 1
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(wq * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t n = 0;  // running key-block number (same count as the MMA thread's)
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int qt = item % p.nqt + p.s_tiles, bh = item / p.nqt;
      const int h = bh % p.heads, b = bh / p.heads;
      const int nkb = item_nkb(qt);
      const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.f;
      const bool plain = !p.causal && slope == 0.f;
      const bool live = qt * QT + wq * 32 < p.T;  // warps whose rows all lie beyond T only keep the barrier protocol going
      const float slope2 = slope * LOG2E;
      const int qidx = qt * QT + row;
      float l = 0.f;
      float m_run = -INFINITY;  // log2 domain
      for (int j = 0; j < nkb; ++j, ++n) {
        const uint32_t buf = n % NBUF;
        const uint32_t tmem_s = tmem_base + lane_addr + buf * KT;
        mbar_wait(&s_full[buf], (n / NBUF) & 1);
        tc_fence_after();
        const int valid = p.T - j * KT;                                   // keys of this block that exist
        const int vrow = p.causal ? min(valid, qidx - j * KT + 1) : valid;  // ... and that this row may attend to (may be <= 0)
        const float bias0 = slope2 * static_cast<float>(j * KT);
        const int ncols = live ? nkeys(j) : 0;  // columns the PV MMA will read
        uint32_t r0[32], r1[32];  // the two 32-column halves of this row's S block (separate arrays: both must stay in registers)
        tmem_ld_32x32b_x32(tmem_s, r0);  // unconditional: columns past nkeys(j) hold stale data that the ncols tests below skip
        tmem_ld_32x32b_x32(tmem_s + 32, r1);
        tmem_ld_wait();
        if (live) {
          // (1) scores to the log2 domain (+ ALiBi, masks) and the block maximum of this row
          float mblk = -INFINITY;
          float mulc[2] = {1.f, 1.f};  // factor that takes r[] of chunk c to the log2 domain inside step (3)'s FFMA2
          auto to_log2 = [&](uint32_t (&r)[32], const int c) {
            if (c * 32 >= ncols) return;
            if (plain && valid - c * 32 >= 32) {
              // raw scores stay in r[]: max(s) * log2(e) = max(s * log2(e)) (monotone rounding), the scale itself moves to step (3)
              float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
#pragma unroll
                for (int k = 0; k < 4; ++k) m4[k] = fmax3(m4[k], __uint_as_float(r[i + 2 * k]), __uint_as_float(r[i + 2 * k + 1]));
              }
              mblk = fmaxf(mblk, fmax3(fmaxf(m4[0], m4[1]), m4[2], m4[3]) * LOG2E);
              mulc[c] = LOG2E;
            } else if (vrow - c * 32 >= 32) {  // every column of this chunk visible to this row: no masks
              float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float t = fmaf(__uint_as_float(r[i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                r[i] = __float_as_uint(t);
                m4[i & 3] = fmaxf(m4[i & 3], t);
              }
              mblk = fmaxf(mblk, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float t = fmaf(__uint_as_float(r[i]), LOG2E, fmaf(slope2, static_cast<float>(c * 32 + i), bias0));
                t = (c * 32 + i < vrow) ? t : -INFINITY;
                r[i] = __float_as_uint(t);
                mblk = fmaxf(mblk, t);
              }
            }
          };
          to_log2(r0, 0);
          to_log2(r1, 1);
          // (2) lazy rescale: raise the running maximum only when the block exceeds it by more than 2^8 (P stays <= 2^8 in fp16)
          float scale = 1.f;
          if (mblk > m_run + 8.f) {
            scale = ex2a3(m_run - mblk);  // 0 for the first block (m_run = -inf)
            m_run = mblk;
          }
          const float mref = (m_run == -INFINITY) ? 0.f : m_run;  // a row that has seen no key yet: every t is -inf -> P = 0
          // (3) P = 2^(t - m_run), row sum
          float lsum = 0.f;
          // steps (3) and (5) of a 32-column chunk in ONE basic block — exponent FFMA2, ex2, row sum, fp16 hi / lo pairs and the
          // tcgen05.st — so that the scheduler can issue the packing of earlier pairs between the MUFU.EX2 instructions (a warp-wide
          // MUFU holds the SFU for 8 cycles; with the two steps in separate blocks 47 % of the softmax samples were fixed-latency
          // waits behind 32 back-to-back MUFUs, the F2FP / HADD2 / FADD2 work of step (5) queued behind them).
          auto exp_store = [&](uint32_t (&r)[32], const int c) {
            if (c * 32 >= ncols) return;
            const uint64_t mul2 = f2_pack(mulc[c], mulc[c]), neg2 = f2_pack(-mref, -mref);
            uint64_t l2[4] = {0ull, 0ull, 0ull, 0ull};
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              float d0, d1;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(r[2 * u]), __uint_as_float(r[2 * u + 1])), mul2, neg2), d0, d1);
              const float e0 = ex2a3(d0), e1 = ex2a3(d1);
              const uint64_t e = f2_pack(e0, e1);
              l2[u & 3] = f2_add(l2[u & 3], e);
              hi[u] = cvt2h(e0, e1);
              if (NP == 2) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
                float q0, q1;
                f2_unpack(f2_sub(e, f2_pack(hf.x, hf.y)), q0, q1);
                lo[u] = cvt2h(q0, q1);
              }
            }
            float a0, a1;
            f2_unpack(f2_add(f2_add(l2[0], l2[1]), f2_add(l2[2], l2[3])), a0, a1);
            lsum += a0 + a1;
            tmem_st_32x32b_x16(tmem_s + c * 16, hi);
            if (NP == 2) tmem_st_32x32b_x16(tmem_s + 32 + c * 16, lo);
          };
          exp_store(r0, 0);
          exp_store(r1, 1);
          l = fmaf(l, scale, lsum);
          // (4) rare: the running maximum moved -> bring the O accumulated so far to the new reference. PV(n-1) (and with it every
          //     earlier PV) must have completed; PV(n) cannot start before this warp arrives on p_full below.
          if (j > 0 && __any_sync(0xffffffffu, scale != 1.f)) {
            mbar_wait(&pv_done[(n - 1) % NBUF], ((n - 1) / NBUF) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * scale);
              tmem_st_32x32b_x32(tmem_base + lane_addr + O_COL + c * 32, o);
            }
          }
          // (5) the P stores were issued by exp_store; PV(n) may read them once they have landed
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      }
      // ---- epilogue: O / l -> fp16 hi [+ lo | + e4m3 planes]; this thread owns the 64 head-dim columns of its row ----
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const bool wr = live && qidx < p.T;
      const float rl = 1.f / l;
      long long orow_idx = static_cast<long long>(b) * p.Tq + (qidx - p.s_tiles * QT);
      if (p.perm_C > 0) orow_idx = (static_cast<long long>(b / p.perm_C) * p.Tq + qidx) * p.perm_C + b % p.perm_C;
      __half* orow = p.out + orow_idx * p.ldo + h * 64;
      uint8_t* f8 = reinterpret_cast<uint8_t*>(p.out + orow_idx * p.ldo + p.out_lo_off) + h * 64;
      const float sh = p.out_scale, sl = p.out_scale * 2048.f;
      // delta-operand mode (DELTA): this row's exact value also goes to cout[b] when it is the masked row; base row to subtract
      const bool crow = DELTA && wr && p.mask_pos != nullptr && qidx == __ldg(p.mask_pos + b);
      const __half* brow = (DELTA && p.base_o != nullptr) ? p.base_o + static_cast<long long>(qidx) * p.d + h * 64 : nullptr;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {  // 32 head-dim columns at a time (register budget: 128 per thread at two CTAs per SM)
        uint32_t o[32];
        if (live) {
          tmem_ld_32x32b_x32(tmem_base + lane_addr + O_COL + c * 32, o);
          tmem_ld_wait();
        }
        if (c == 1) {  // both halves of O are in registers / already written: the next item's PV(0) may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
        if (wr) {
          uint32_t hi[16];
          uint4* d4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float x0, x1;
            f2_unpack(f2_mul(f2_pack(__uint_as_float(o[2 * u]), __uint_as_float(o[2 * u + 1])), f2_pack(rl, rl)), x0, x1);
            if (DELTA && crow) {  // rare (one row per sequence and head): fp16 hi / lo pair of the full value
              __half* cr = p.cout + static_cast<long long>(b) * p.ldc + h * 64 + c * 32 + 2 * u;
              const uint32_t ch = cvt2h(x0, x1);
              const float2 cf = __half22float2(*reinterpret_cast<const __half2*>(&ch));
              *reinterpret_cast<uint32_t*>(cr) = ch;
              *reinterpret_cast<uint32_t*>(cr + p.c_lo_off) = cvt2h(x0 - cf.x, x1 - cf.y);
            }
            if (DELTA && brow != nullptr) {
              const float2 bq = __half22float2(__ldg(reinterpret_cast<const __half2*>(brow + c * 32 + 2 * u)));
              x0 -= bq.x;
              x1 -= bq.y;
            }
            hi[u] = cvt2h(x0, x1);
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[u]));
            float q0, q1;
            f2_unpack(f2_sub(f2_pack(x0, x1), f2_pack(hf.x, hf.y)), q0, q1);
            o[2 * u] = __float_as_uint(q0);           // o[] now holds the fp32 remainders
            o[2 * u + 1] = __float_as_uint(q1);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) d4[u] = make_uint4(hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
          if (p.out_fmt == 1) {
            uint4* l4 = reinterpret_cast<uint4*>(orow + p.out_lo_off + c * 32);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              l4[u] = make_uint4(cvt2h(__uint_as_float(o[8 * u]), __uint_as_float(o[8 * u + 1])),
                                 cvt2h(__uint_as_float(o[8 * u + 2]), __uint_as_float(o[8 * u + 3])),
                                 cvt2h(__uint_as_float(o[8 * u + 4]), __uint_as_float(o[8 * u + 5])),
                                 cvt2h(__uint_as_float(o[8 * u + 6]), __uint_as_float(o[8 * u + 7])));
          } else if (p.out_fmt == 2) {
            uint4* l4 = reinterpret_cast<uint4*>(f8 + c * 32);
            uint4* h4 = reinterpret_cast<uint4*>(f8 + p.d + c * 32);
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              uint32_t wl[4], wh[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int e = 16 * v + 4 * u;  // first of four consecutive columns
                const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hi[e / 2]));
                const float2 h23 = __half22float2(*reinterpret_cast<const __half2*>(&hi[e / 2 + 1]));
                wh[u] = pack4_e4m3(h01.x * sh, h01.y * sh, h23.x * sh, h23.y * sh);
                wl[u] = pack4_e4m3(__uint_as_float(o[e]) * sl, __uint_as_float(o[e + 1]) * sl, __uint_as_float(o[e + 2]) * sl,
                                   __uint_as_float(o[e + 3]) * sl);
              }
              l4[v] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
              h4[v] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_attention_tc4(const AttnLaunch& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return PG_OK;
  if (a.nseg != 1 && a.nseg != 3) return set_error(PG_ERR_ARG, "attention: nseg must be 1 or 3");
  if (a.ld % 8 || a.lo_off % 8 || a.ldo % 8 || a.out_lo_off % 8 || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return set_error(PG_ERR_ARG, "attention_tc4: pitches must be multiples of 8 elements and out 16-byte aligned");
  if (a.q_begin != 0) return set_error(PG_ERR_ARG, "attention_tc4: q_begin is only supported by the mma.sync kernel");
  if (a.prefix && (!a.causal || a.prefix_len <= 0 || a.prefix_len % QT))
    return set_error(PG_ERR_ARG, "attention_tc4: a shared prefix needs causal attention and a length that is a multiple of 128");
  Attn4Params p{};
  p.B = a.B; p.heads = a.heads; p.d = a.heads * 64;
  p.Tq = a.T; p.s_tiles = a.prefix ? a.prefix_len / QT : 0; p.s_blocks = a.prefix ? a.prefix_len / KT : 0;
  p.T = a.T + p.s_tiles * QT;
  p.nqt = (a.T + QT - 1) / QT; p.nkb = (p.T + KT - 1) / KT;
  p.lo_off = a.lo_off; p.out = a.out; p.ldo = a.ldo; p.out_lo_off = a.out_lo_off;
  p.causal = a.causal; p.alibi_slopes = a.alibi_slopes;
  p.out_fmt = a.out_fmt < 0 ? (a.out_lo_off > 0 ? 1 : 0) : a.out_fmt;
  p.out_scale = a.out_scale;
  p.perm_C = a.perm_C;
  p.base_o = a.base_o; p.mask_pos = a.mask_pos; p.cout = a.cout; p.ldc = a.ldc; p.c_lo_off = a.c_lo_off;
  if ((a.base_o || a.mask_pos) && (a.prefix || a.perm_C || (a.mask_pos && (!a.cout || a.ldc % 2 || a.c_lo_off % 2 || a.c_lo_off <= 0))))
    return set_error(PG_ERR_ARG, "attention_tc4: delta-operand mode needs plain sequences and, with mask_pos, a compact hi/lo output");
  if (a.perm_C < 0 || (a.perm_C > 0 && (a.prefix || a.B % a.perm_C))) return set_error(PG_ERR_ARG, "attention_tc4: bad column-attention arguments");
  if (p.out_fmt > 2 || (p.out_fmt >= 1 && a.out_lo_off <= 0) || (p.out_fmt == 2 && !(a.out_scale > 0.f)))
    return set_error(PG_ERR_ARG, "attention_tc4: bad output format");
  const int np = a.nseg == 3 ? 2 : 1;
  const uint64_t width = static_cast<uint64_t>(3) * p.d * np;
  if (np == 2 && a.lo_off != 3ll * p.d) return set_error(PG_ERR_ARG, "attention_tc4: lo planes must follow the hi planes (lo_off == 3*d)");
  CUtensorMap tmQ, tm, tmP;
  int rc = make_tmap_f16_2d(&tmQ, a.qkv, static_cast<uint64_t>(a.B) * a.T, width, a.ld, QT, 64);
  if (rc) return rc;
  rc = make_tmap_f16_2d(&tm, a.qkv, static_cast<uint64_t>(a.B) * a.T, width, a.ld, KT, 64);
  if (rc) return rc;
  tmP = tm;
  if (a.prefix) {
    rc = make_tmap_f16_2d(&tmP, a.prefix, static_cast<uint64_t>(a.prefix_len), width, a.ld, KT, 64);
    if (rc) return rc;
  }
  const long long nitems = static_cast<long long>(a.B) * a.heads * p.nqt;
  const int cap = 2 * num_sms();
  const int grid = nitems < cap ? static_cast<int>(nitems) : cap;
  int dev = 0;
  PG_CUDA_OK(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (dev < 64 && !attr_set[dev]) {
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc4_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem4<1>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc4_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem4<2>::TOTAL));
    PG_CUDA_OK(cudaFuncSetAttribute(attn_tc4_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem4<2>::TOTAL));
    attr_set[dev] = true;
  }
  const bool delta = a.base_o != nullptr || a.mask_pos != nullptr;
  if (delta && np != 2) return set_error(PG_ERR_UNSUPPORTED, "attention_tc4: the delta-operand form runs on fp16 hi/lo operands (nseg 3)");
  if (delta) attn_tc4_kernel<2, 1><<<grid, ATT_THREADS, Smem4<2>::TOTAL, s>>>(tmQ, tm, tmP, p);
  else if (np == 1) attn_tc4_kernel<1, 0><<<grid, ATT_THREADS, Smem4<1>::TOTAL, s>>>(tmQ, tm, tmP, p);
  else attn_tc4_kernel<2, 0><<<grid, ATT_THREADS, Smem4<2>::TOTAL, s>>>(tmQ, tm, tmP, p);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
