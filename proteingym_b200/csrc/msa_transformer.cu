// Alignment-specific kernels of the MSA Transformer forward (esm/model/msa_transformer.py:150-222, esm/axial_attention.py):
// the embedding of a [B, R, C] block of masked alignments, and the data movement + softmax around the two tensor-core GEMMs of the
// TIED row attention (axial_attention.py:115-181):
//     S[b,h][i,j] = sum_r sum_d q[b,r,i,h,d] k[b,r,j,h,d]          one [C, C] map per head, K = R * 64   (grouped GEMM, gemm_tc.cu)
//     P           = softmax_j(S / sqrt(R))                          (q carries head_dim^-1/2 already)
//     ctx[b,r,i,h,:] = sum_j P[b,h][i,j] v[b,r,j,h,:]               N = R * 64, K = C                     (grouped GEMM)
// Both products are plain "A x W^T" GEMMs once q / k are regrouped as [(b,h,i), (r,d)], v as [(b,h,r,d), j] and the result is
// scattered back to token rows; those regroupings are the 128-byte-chunk permutations and the 64x64 transpose below (HBM-bound
// copies, ~5 % of a layer). The column attention needs nothing here: LayerNorm writes its rows in (b, c, r) order and the
// attention kernel writes its output back in (b, r, c) order (attention_tc4.cu perm_C).
#include "common.h"
#include "ptx.cuh"

namespace pg {
namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < nw) ? sh[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// One block per token row (b, r, c): x = LayerNorm_before(E[tok] + Pos[c + 2] + RowPos[r]); the token at (row 0, column positions[p])
// is <mask>. msa_transformer.py:157-172; LearnedPositionalEmbedding positions = column + 2 (modules.py:254-271), no padding here.
constexpr int EMB_MAXV = 8;  // d <= 256 * 8
__global__ void msa_embed_kernel(MsaEmbedLaunch e) {
  __shared__ float sh[32];
  const long long row = blockIdx.x;
  const int c = static_cast<int>(row % e.Cw);
  const int r = static_cast<int>((row / e.Cw) % e.R);
  const int b = static_cast<int>(row / (static_cast<long long>(e.Cw) * e.R));
  const int p = e.p_offset + b;
  const int col = (e.win_start ? e.win_start[p] : 0) + c;
  int tok = e.tokens[static_cast<long long>(r) * e.Cfull + col];
  if (r == 0 && col == e.positions[p]) tok = e.mask_idx;
  const float* et = e.embed + static_cast<long long>(tok) * e.d;
  const float* pt = e.pos_table + static_cast<long long>(c + 2) * e.d;
  const float* rt = e.row_pos ? e.row_pos + static_cast<long long>(r) * e.d : nullptr;
  float v[EMB_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < EMB_MAXV; ++i) {
    const int k = threadIdx.x + i * blockDim.x;
    v[i] = 0.f;
    if (k < e.d) {
      v[i] = et[k] + pt[k];
      if (rt) v[i] += rt[k];
      s += v[i];
    }
  }
  const float mean = block_sum(s, sh) / e.d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < EMB_MAXV; ++i) {
    const int k = threadIdx.x + i * blockDim.x;
    if (k < e.d) ss += (v[i] - mean) * (v[i] - mean);
  }
  const float rstd = rsqrtf(block_sum(ss, sh) / e.d + 1e-5f);
  float* xr = e.x + row * e.d;
#pragma unroll
  for (int i = 0; i < EMB_MAXV; ++i) {
    const int k = threadIdx.x + i * blockDim.x;
    if (k < e.d) xr[k] = (v[i] - mean) * rstd * e.gamma[k] + e.beta[k];
  }
}

// q and k of the fused projection, token rows (b, r, i) x [plane][which][h][64] -> tied rows (b, h, i) x [plane][r][64].
// One 16-byte unit per thread, 8 threads per 128-byte chunk; r is the fastest chunk index so a tied row is written contiguously.
__global__ void tied_gather_qk_kernel(const __half* __restrict__ qkv, long long ldq, long long lo_off, int B, int R, int C, int H, int np,
                                      int Cp, __half* __restrict__ tq, __half* __restrict__ tk, long long ldt) {
  const long long unit = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int u = static_cast<int>(unit & 7);
  long long ch = unit >> 3;
  const long long total = 2ll * B * H * C * np * R;
  if (ch >= total) return;
  const int r = static_cast<int>(ch % R); ch /= R;
  const int pl = static_cast<int>(ch % np); ch /= np;
  const int i = static_cast<int>(ch % C); ch /= C;
  const int h = static_cast<int>(ch % H); ch /= H;
  const int b = static_cast<int>(ch % B); ch /= B;
  const int which = static_cast<int>(ch);  // 0 = q, 1 = k
  const int d = H * 64;
  const __half* src = qkv + (static_cast<long long>(b) * R + r) * C * ldq + static_cast<long long>(i) * ldq + pl * lo_off + which * d + h * 64;
  __half* dst = (which ? tk : tq) + ((static_cast<long long>(b) * H + h) * Cp + i) * ldt + static_cast<long long>(pl) * R * 64 + r * 64;
  reinterpret_cast<uint4*>(dst)[u] = reinterpret_cast<const uint4*>(src)[u];
}

// v, token rows (b, r, j) x [plane][2d + h*64 + dd] -> tied rows (b, h, r, dd) x [plane][j] (j padded with zeros to Kp): 64x64 tiles
// through shared memory.
__global__ void tied_transpose_v_kernel(const __half* __restrict__ qkv, long long ldq, long long lo_off, int R, int C, int H, int Kp,
                                        __half* __restrict__ tv, long long ldv) {
  __shared__ __half tile[64][72];
  const int jt = blockIdx.x, r = blockIdx.y;
  const int h = blockIdx.z % H;
  const int bp = blockIdx.z / H;  // b * np + plane
  const int np = (lo_off > 0) ? 2 : 1;
  const int b = bp / np, pl = bp % np;
  const int d = H * 64;
  const int t = threadIdx.x;
  {
    const int j = jt * 64 + (t >> 2), d0 = (t & 3) * 16;
    uint4 a = make_uint4(0, 0, 0, 0), c = a;
    if (j < C) {
      const __half* src = qkv + ((static_cast<long long>(b) * R + r) * C + j) * ldq + pl * lo_off + 2 * d + h * 64 + d0;
      a = reinterpret_cast<const uint4*>(src)[0];
      c = reinterpret_cast<const uint4*>(src)[1];
    }
    *reinterpret_cast<uint4*>(&tile[t >> 2][d0]) = a;
    *reinterpret_cast<uint4*>(&tile[t >> 2][d0 + 8]) = c;
  }
  __syncthreads();
  {
    const int dd = t >> 2, j0 = (t & 3) * 16;
    __half out[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) out[k] = tile[j0 + k][dd];
    __half* dst = tv + (((static_cast<long long>(b) * H + h) * R + r) * 64 + dd) * ldv + static_cast<long long>(pl) * Kp + jt * 64 + j0;
    reinterpret_cast<uint4*>(dst)[0] = *reinterpret_cast<const uint4*>(&out[0]);
    reinterpret_cast<uint4*>(dst)[1] = *reinterpret_cast<const uint4*>(&out[8]);
  }
}

// One warp per attention row (b, h, i): P = softmax_j(scale * S[., j]), j < C, as fp16 hi [+ lo] planes of Kp columns (zero padded).
__global__ void tied_softmax_kernel(const float* __restrict__ S, long long lds, int G, int C, int Cp, int Kp, float scale,
                                    __half* __restrict__ P, long long ldp, int np) {
  const long long w = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<long long>(G) * C) return;
  const long long row = (w / C) * Cp + (w % C);
  const float* s = S + row * lds;
  float m = -INFINITY;
  for (int j = lane; j < C; j += 32) m = fmaxf(m, s[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
  for (int j = lane; j < C; j += 32) l += expf(s[j] * scale - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  __half* p = P + row * ldp;
  for (int j = lane; j < Kp; j += 32) {
    const float v = (j < C) ? expf(s[j] * scale - m) / l : 0.f;
    __half hi, lo;
    split_hi_lo(v, hi, lo);
    p[j] = hi;
    if (np == 2) p[Kp + j] = lo;
  }
}

// Context rows back to token order: for every (b, h, i, r) the 64-column chunk r of tied row (b, h, i) goes to columns h*64.. of token
// row (b, r, i), plane by plane (`cb` bytes per chunk: 128 for fp16 planes, 64 for e4m3 planes). h is the fastest chunk index so a token
// row is written contiguously.
__global__ void tied_scatter_out_kernel(const uint8_t* __restrict__ src, long long src_pitch, long long src_plane, uint8_t* __restrict__ dst,
                                        long long dst_pitch, long long dst_plane, int cb, int B, int R, int C, int H, int Cp) {
  const int upc = cb >> 4;  // 16-byte units per chunk
  const long long unit = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int u = static_cast<int>(unit % upc);
  long long ch = unit / upc;
  if (ch >= static_cast<long long>(B) * R * C * H) return;
  const int h = static_cast<int>(ch % H); ch /= H;
  const int i = static_cast<int>(ch % C); ch /= C;
  const int r = static_cast<int>(ch % R); ch /= R;
  const int b = static_cast<int>(ch);
  const uint8_t* s = src + ((static_cast<long long>(b) * H + h) * Cp + i) * src_pitch + src_plane + static_cast<long long>(r) * cb;
  uint8_t* d = dst + ((static_cast<long long>(b) * R + r) * C + i) * dst_pitch + dst_plane + static_cast<long long>(h) * cb;
  reinterpret_cast<uint4*>(d)[u] = reinterpret_cast<const uint4*>(s)[u];
}

}  // namespace

int launch_msa_embed(const MsaEmbedLaunch& e, cudaStream_t s) {
  const long long rows = static_cast<long long>(e.B) * e.R * e.Cw;
  if (rows <= 0) return PG_OK;
  if (e.d > 256 * EMB_MAXV) return set_error(PG_ERR_UNSUPPORTED, "msa embed: embed_dim > 2048");
  msa_embed_kernel<<<static_cast<unsigned>(rows), 256, 0, s>>>(e);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

int launch_tied_gather_qk(const __half* qkv, int64_t ldq, int64_t lo_off, int B, int R, int C, int H, int Cp, __half* tq, __half* tk,
                          int64_t ldt, cudaStream_t s) {
  const int np = lo_off > 0 ? 2 : 1;
  const long long units = 2ll * B * H * C * np * R * 8;
  tied_gather_qk_kernel<<<static_cast<unsigned>((units + 255) / 256), 256, 0, s>>>(qkv, ldq, lo_off, B, R, C, H, np, Cp, tq, tk, ldt);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

int launch_tied_transpose_v(const __half* qkv, int64_t ldq, int64_t lo_off, int B, int R, int C, int H, int Kp, __half* tv, int64_t ldv,
                            cudaStream_t s) {
  const int np = lo_off > 0 ? 2 : 1;
  if (R > 65535 || static_cast<long long>(B) * np * H > 65535) return set_error(PG_ERR_ARG, "tied attention: grid too large");
  dim3 grid(Kp / 64, R, B * np * H);
  tied_transpose_v_kernel<<<grid, 256, 0, s>>>(qkv, ldq, lo_off, R, C, H, Kp, tv, ldv);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

int launch_tied_softmax(const float* S, int64_t lds, int G, int C, int Cp, int Kp, float scale, __half* P, int64_t ldp, int np, cudaStream_t s) {
  const long long rows = static_cast<long long>(G) * C;
  tied_softmax_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(S, lds, G, C, Cp, Kp, scale, P, ldp, np);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

// fmt: operand format of both buffers (common.h); tied rows hold N = R * 64 columns per plane, token rows d = H * 64.
int launch_tied_scatter_out(const __half* ot, int64_t ldo_t, __half* out, int64_t ldo, int fmt, int B, int R, int C, int H, int Cp,
                            cudaStream_t s) {
  const long long Nt = static_cast<long long>(R) * 64, d = static_cast<long long>(H) * 64;
  const long long chunks = static_cast<long long>(B) * R * C * H;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(ot);
  uint8_t* dst = reinterpret_cast<uint8_t*>(out);
  auto run = [&](long long sp, long long dp, int cb) {
    const long long units = chunks * (cb >> 4);
    tied_scatter_out_kernel<<<static_cast<unsigned>((units + 255) / 256), 256, 0, s>>>(src, ldo_t * 2, sp, dst, ldo * 2, dp, cb, B, R, C, H, Cp);
  };
  run(0, 0, 128);
  if (fmt == 1) run(2 * Nt, 2 * d, 128);
  if (fmt == 2) { run(2 * Nt, 2 * d, 64); run(3 * Nt, 3 * d, 64); }
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // namespace pg
