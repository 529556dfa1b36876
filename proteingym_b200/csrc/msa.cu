// MSA pre-processing kernels for the Tranception retrieval prior — integer / byte work, HBM- and
// ALU-bound, nothing here belongs on tensor cores:
//   msa_cluster_kernel   sequence re-weighting: for every sequence i the number of sequences j whose identity to i, counted
//                        over i's non-gap positions, exceeds the threshold (reference: proteingym/utils/weights.py:164-216,
//                        calc_num_cluster_members_nogaps_parallel; O(N^2 L) byte compares, numba on CPU in the reference)
//   msa_prior_kernel     weighted amino-acid frequencies per alignment column with a 1e-5 pseudocount
//                        (reference: proteingym/baselines/tranception/tranception/utils/msa_utils.py:118-128)
#include "common.h"

namespace pg {

namespace {

constexpr int TI = 64, TJ = 64, LC = 256;  // i-tile, j-tile (sequences), column chunk (bytes)

// tokens [N, ld] uint8, 0 = gap / invalid. min_matches[i] = smallest match count that puts j in i's cluster (host computes it with
// the reference's exact float64 expression pair_matches / L_non_gaps[i] > identity_threshold). neighbors[i] = 1 + #{j != i: ...}.
__global__ void __launch_bounds__(256) msa_cluster_kernel(const uint8_t* __restrict__ tok, long long ld, int N, int L,
                                                          const int32_t* __restrict__ min_matches, int32_t* __restrict__ neighbors) {
  __shared__ uint32_t si[TI][LC / 4 + 1];
  __shared__ uint32_t sj[TJ][LC / 4 + 1];
  __shared__ int cnt[TI];
  const int i0 = blockIdx.x * TI;
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;  // 4x4 (i, j) pairs per thread
  if (threadIdx.x < TI) cnt[threadIdx.x] = 0;
  int need[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) need[a] = (i0 + ti + a < N) ? min_matches[i0 + ti + a] : 0x7fffffff;
  for (int j0 = 0; j0 < N; j0 += TJ) {
    int m[4][4] = {};
    for (int c0 = 0; c0 < L; c0 += LC) {
      __syncthreads();
      for (int w = threadIdx.x; w < TI * (LC / 4); w += 256) {  // cooperative, coalesced 4-byte loads; zero beyond N / L
        const int r = w / (LC / 4), cw = w % (LC / 4);
        const int col = c0 + cw * 4;
        uint32_t vi = 0, vj = 0;
        if (col < L) {
          const int nb = L - col < 4 ? L - col : 4;
          if (i0 + r < N) {
            const uint8_t* p = tok + static_cast<long long>(i0 + r) * ld + col;
            for (int b = 0; b < nb; ++b) vi |= static_cast<uint32_t>(p[b]) << (8 * b);
          }
          if (j0 + r < N) {
            const uint8_t* p = tok + static_cast<long long>(j0 + r) * ld + col;
            for (int b = 0; b < nb; ++b) vj |= static_cast<uint32_t>(p[b]) << (8 * b);
          }
        }
        si[r][cw] = vi;
        sj[r][cw] = vj;
      }
      __syncthreads();
#pragma unroll 4
      for (int w = 0; w < LC / 4; ++w) {
        uint32_t a[4], nz[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = si[ti + q][w];
          nz[q] = __vcmpne4(a[q], 0u);  // 0xff in every byte of i that is not a gap
          b[q] = sj[tj + q][w];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int r = 0; r < 4; ++r) m[q][r] += __popc(__vcmpeq4(a[q], b[r]) & nz[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int c = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int gi = i0 + ti + q, gj = j0 + tj + r;
        if (gi < N && gj < N && gi != gj && (m[q][r] >> 3) >= need[q]) ++c;  // popc counts 8 bits per matching byte
      }
      if (c) atomicAdd(&cnt[ti + q], c);
    }
  }
  __syncthreads();
  if (threadIdx.x < TI && i0 + threadIdx.x < N) neighbors[i0 + threadIdx.x] = 1 + cnt[threadIdx.x];
}

// tokens_t [L, N] uint8 (token id < vocab, anything >= vocab = not in vocabulary), weights [N] fp64.
// out[j, k] = (sum_i w_i [tok_ij == k] + base * W) / (sum_i w_i [tok_ij in vocab] + vocab * base * W),  W = sum_i w_i.
template <int V>
__global__ void __launch_bounds__(128) msa_prior_kernel(const uint8_t* __restrict__ tok_t, const double* __restrict__ w, int N,
                                                        double base, double* __restrict__ out) {
  // Each thread keeps its V partial sums in its own column of shared memory and adds w_i to ONE of them per sequence (the first
  // version ran V predicated fp64 adds per sequence and was fp64-ALU bound at a sixth of the HBM roofline). Order of additions per
  // thread and the reduction tree are fixed: the result does not depend on scheduling.
  __shared__ double sb[V][128];
  __shared__ double red[128];
  const int j = blockIdx.x, tid = threadIdx.x;
  const uint8_t* col = tok_t + static_cast<long long>(j) * N;
#pragma unroll
  for (int k = 0; k < V; ++k) sb[k][tid] = 0.0;
  double wsum = 0.0;
  for (int i = tid; i < N; i += 128) {
    const int t = col[i];
    const double wi = w[i];
    wsum += wi;
    if (t < V) sb[t][tid] += wi;
  }
  auto block_reduce = [&](double v) {  // fixed-order tree: deterministic
    __syncthreads();
    red[tid] = v;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
      if (tid < s) red[tid] += red[tid + s];
      __syncthreads();
    }
    return red[0];
  };
  const double W = block_reduce(wsum);
  double tot[V];
  double norm = 0.0;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    tot[k] = block_reduce(sb[k][tid]) + base * W;
    norm += tot[k];
  }
  if (tid == 0)
    for (int k = 0; k < V; ++k) out[static_cast<long long>(j) * V + k] = tot[k] / norm;
}

// EVE Bayesian decoder, output 1x1 convolution with PER-SAMPLE weights (batched local-reparameterisation sampler, eve_prior.py;
// reference arithmetic: trancepteve/EVE/VAE_decoder.py:139-147 applied to sample s's own draw of the convolution weights):
//   y[s, j*C + c] = sum_a x[s, j*A + a] * conv[s, c*A + a]        x [S, J*A], conv [S, C*A], y [S, J*C]
// One block per sample: its x row and its C*A weights are staged in shared memory once, then every thread owns (j, c) outputs.
__global__ void __launch_bounds__(256) eve_conv_kernel(const float* __restrict__ x, const float* __restrict__ conv, int J, int A, int C,
                                                        float* __restrict__ y) {
  extern __shared__ float sm[];
  float* sx = sm;            // [J*A]
  float* sc = sm + J * A;    // [C*A]
  const long long s = blockIdx.x;
  for (int i = threadIdx.x; i < J * A; i += blockDim.x) sx[i] = x[s * J * A + i];
  for (int i = threadIdx.x; i < C * A; i += blockDim.x) sc[i] = conv[s * C * A + i];
  __syncthreads();
  for (int o = threadIdx.x; o < J * C; o += blockDim.x) {
    const int j = o / C, c = o - j * C;
    float acc = 0.f;
    for (int a = 0; a < A; ++a) acc = fmaf(sx[j * A + a], sc[c * A + a], acc);  // fixed order: deterministic
    y[s * J * C + o] = acc;
  }
}

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int pg_eve_output_conv(const float* x, const float* conv, int32_t S, int32_t J, int32_t A, int32_t C, float* y, pg_stream stream) {
  if (S < 0 || J <= 0 || A <= 0 || C <= 0) return set_error(PG_ERR_ARG, "pg_eve_output_conv: bad sizes");
  if (S == 0) return PG_OK;
  if (!x || !conv || !y) return set_error(PG_ERR_ARG, "pg_eve_output_conv: null buffer");
  const size_t smem = (static_cast<size_t>(J) * A + static_cast<size_t>(C) * A) * sizeof(float);
  if (smem > 48 * 1024) return set_error(PG_ERR_UNSUPPORTED, "pg_eve_output_conv: (J + C) * A floats must fit 48 KB of shared memory");
  eve_conv_kernel<<<S, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, conv, J, A, C, y);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

int pg_msa_cluster_neighbors(const uint8_t* tokens, int64_t ld, int32_t N, int32_t L, const int32_t* min_matches,
                             int32_t* out_neighbors, pg_stream stream) {
  if (N < 0 || L < 0 || ld < L) return set_error(PG_ERR_ARG, "pg_msa_cluster_neighbors: bad sizes");
  if (N == 0) return PG_OK;
  if (!tokens || !min_matches || !out_neighbors) return set_error(PG_ERR_ARG, "pg_msa_cluster_neighbors: null buffer");
  msa_cluster_kernel<<<(N + TI - 1) / TI, 256, 0, static_cast<cudaStream_t>(stream)>>>(tokens, ld, N, L, min_matches, out_neighbors);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

int pg_msa_prior(const uint8_t* tokens_t, const double* weights, int32_t N, int32_t L, int32_t vocab, double base_rate, double* out,
                 pg_stream stream) {
  if (N <= 0 || L < 0) return set_error(PG_ERR_ARG, "pg_msa_prior: bad sizes");
  if (vocab != 25) return set_error(PG_ERR_UNSUPPORTED, "pg_msa_prior: vocab must be 25 (Tranception tokenizer)");
  if (L == 0) return PG_OK;
  if (!tokens_t || !weights || !out) return set_error(PG_ERR_ARG, "pg_msa_prior: null buffer");
  msa_prior_kernel<25><<<L, 128, 0, static_cast<cudaStream_t>(stream)>>>(tokens_t, weights, N, base_rate, out);
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

}  // extern "C"
