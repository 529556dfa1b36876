// C-ABI of libpgscore.so (include/pgscore.h): model handle, weight repacking, and the batched masked-marginal forward.
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace pg {

std::string& tls_error() {
  static thread_local std::string e;
  return e;
}
int set_error(int code, const std::string& msg) {
  tls_error() = msg;
  return code;
}
int launch_attention_tc(const AttnLaunch& a, cudaStream_t s) {
  return launch_attention_tc4(a, s);
}

int num_sms() {  // of the current device (cached per ordinal: one process may hold handles on several GPUs)
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// ---- launch accounting / profiling -------------------------------------------------------------------------------
namespace {
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;      // pairs
  std::vector<int> ev_cat;
  size_t used = 0;
  long long launches[CAT_COUNT] = {};
};
ProfState& prof() {
  static ProfState p;
  return p;
}
std::mutex& prof_mu() {
  static std::mutex m;
  return m;
}
}  // namespace

ProfScope::ProfScope(int c, cudaStream_t st, int launches) : cat(c), s(st), slot(-1) {
  std::lock_guard<std::mutex> lk(prof_mu());
  ProfState& p = prof();
  p.launches[c] += launches;
  if (!p.on) return;
  if (p.used + 2 > p.ev.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      p.ev.push_back(e);
    }
    p.ev_cat.push_back(c);
  }
  slot = static_cast<int>(p.used);
  p.ev_cat[slot / 2] = c;
  p.used += 2;
  cudaEventRecord(p.ev[slot], s);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(prof_mu());
  cudaEventRecord(prof().ev[slot + 1], s);
}

namespace {

// fp32 [N, K] -> fp16 hi at [n, 0:K] and (np == 2) lo at [n, K:2K]; rows scaled by `scale` (q scaling fold).
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int K, float scale, __half* __restrict__ out, int np) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  const float v = w[i] * scale;
  __half hi, lo;
  split_hi_lo(v, hi, lo);
  out[static_cast<long long>(n) * K * np + k] = hi;
  if (np == 2) out[static_cast<long long>(n) * K * np + K + k] = lo;
}
// Conv1D weights are stored [in = K, out = N] (x @ W); repack to the K-major [N, np*K] operand layout.
__global__ void pack_weight_t_kernel(const float* __restrict__ w, int N, int K, __half* __restrict__ out, int np) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  __half hi, lo;
  split_hi_lo(w[static_cast<long long>(k) * N + n], hi, lo);
  out[static_cast<long long>(n) * K * np + k] = hi;
  if (np == 2) out[static_cast<long long>(n) * K * np + K + k] = lo;
}
// Weights in the fp16 + e4m3 format of gemm_tc.cu (nseg 2): row n = [hi fp16 (K) | hi8 (K bytes) | lo8 (K bytes)], pitch 4K bytes,
// hi8 = e4m3(hi * t), lo8 = e4m3(lo * 2^11 * t) with ONE power of two t per matrix: the one that puts the largest |hi| of the matrix in
// (112, 224]. e4m3 is a floating-point format (2^-4 relative precision over 15 binades), so rows much smaller than the largest one
// keep their precision; a per-row scale would only cost the epilogue a load per output column. w_inv[n] = 1 / t for every n.
// Element (n, k) of the source is w[n * sn + k * sk] (sk = 1: nn.Linear; sn = 1: Conv1D). Two kernels: matrix |hi| maximum, then pack.
__global__ void weight_absmax_kernel(const float* __restrict__ w, int N, int K, long long sn, long long sk, float scale,
                                     unsigned int* __restrict__ amax_bits) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float a = 0.f;
  if (i < static_cast<long long>(N) * K) a = fabsf(__half2float(f2h_sat(w[(i / K) * sn + (i % K) * sk] * scale)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
  if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(amax_bits, __float_as_uint(a));  // non-negative floats order like their bit patterns
}
__global__ void pack_weight_f8_kernel(const float* __restrict__ w, int N, int K, long long sn, long long sk, float scale,
                                      const unsigned int* __restrict__ amax_bits, __half* __restrict__ out, float* __restrict__ w_inv) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float amax = __uint_as_float(*amax_bits);
  float t = 1.f;
  if (amax > 0.f) {
    int e;
    frexpf(amax, &e);                 // amax = m * 2^e, m in [0.5, 1)
    t = ldexpf(1.f, 8 - e);           // amax * t in [128, 256)
    if (amax * t > 224.f) t *= 0.5f;  // -> (112, 224]: one binade of headroom below the e4m3 maximum of 448
  }
  const float* wr = w + static_cast<long long>(n) * sn;
  __half* orow = out + static_cast<long long>(n) * K * 2;
  uint8_t* f8 = reinterpret_cast<uint8_t*>(orow + K);
  for (int k = lane; k < K; k += 32) {
    const float v = wr[k * sk] * scale;
    __half hi, lo;
    split_hi_lo(v, hi, lo);
    const float hf = __half2float(hi);
    orow[k] = hi;
    f8[k] = static_cast<uint8_t>(cvt_e4m3x2(hf * t, 0.f) & 0xff);
    f8[K + k] = static_cast<uint8_t>(cvt_e4m3x2((v - hf) * t * 2048.f, 0.f) & 0xff);
  }
  if (lane == 0) w_inv[n] = 1.f / t;
}
// One matrix: absmax pass + pack pass on stream s; `scratch` is one device word per call (kept alive by the caller).
void launch_pack_weight_f8(const float* w, int N, int K, long long sn, long long sk, float scale, unsigned int* scratch, __half* out,
                           float* w_inv, cudaStream_t s) {
  const long long tot = static_cast<long long>(N) * K;
  cudaMemsetAsync(scratch, 0, sizeof(unsigned int), s);
  weight_absmax_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(w, N, K, sn, sk, scale, scratch);
  pack_weight_f8_kernel<<<(N + 7) / 8, 256, 0, s>>>(w, N, K, sn, sk, scale, scratch, out, w_inv);
}
__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * scale;
}

struct Layer {
  __half *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *ln1g = nullptr, *ln1b = nullptr, *ln2g = nullptr, *ln2b = nullptr;
  float *iqkv = nullptr, *io = nullptr, *i1 = nullptr, *i2 = nullptr;  // PG_PREC_F16F8: 1 / t_n per weight row
  float* conv_taps = nullptr;  // Tranception: [3][4][64][8]
};

// Power-of-two scales the producers of A operands apply to their e4m3 planes (PG_PREC_F16F8; gemm_tc.cu header). e4m3 keeps
// ~2^-4 relative precision from 2^-6 to 448 (scaled), so the scale only places the window: values above 448/s saturate (their
// cross-term contribution degrades towards single-fp16 accuracy, nothing worse), values below 2^-9/s vanish from the cross term.
// Measured insensitive between 1 and 16 (scripts/precision_f8.py).
constexpr float S_LN = 4.f;     // LayerNorm output            -> QKV / fc1 GEMM     (|x| up to 112)
constexpr float S_ATT = 4.f;    // attention output            -> out_proj GEMM
constexpr float S_GELU = 2.f;   // GELU(fc1) (ESM)             -> fc2 GEMM           (up to 224)
constexpr float S_RELU2 = 1.f;  // relu(fc1)^2 (Tranception)   -> fc2 GEMM           (up to 448)

}  // namespace
}  // namespace pg

struct pg_handle_s {
  pg_model_desc desc{};
  int np = 1;    // operand row pitch in units of the fp16 hi plane: 1 (fp16) or 2 (hi | lo, or hi | e4m3 planes)
  int nseg = 1;  // GEMM operand mode: 1 fp16, 3 fp16 hi/lo x3, 2 fp16 + e4m3 cross terms
  bool loaded = false;
  std::string err;
  std::vector<void*> allocs;
  std::vector<pg::Layer> layers;
  float *embed = nullptr, *pos = nullptr, *lnbg = nullptr, *lnbb = nullptr, *lnag = nullptr, *lnab = nullptr;
  float *hdw = nullptr, *hdb = nullptr, *hlng = nullptr, *hlnb = nullptr, *hbias = nullptr;
  float *rot_cos = nullptr, *rot_sin = nullptr;
  int rot_rows = 0;
  // workspace
  long long max_rows = 0;
  int head_cap = 0;
  float* x = nullptr;
  __half *abuf = nullptr, *qkv = nullptr, *fbuf = nullptr;
  float *hs_a = nullptr, *hs_b = nullptr;
  int32_t* row_sel = nullptr;  // [head_cap] token index within the window to emit
  // Tranception
  __half* qkv2 = nullptr;
  float *tok_logp = nullptr, *slopes = nullptr;
  // exact wild-type-prefix reuse (pg_ar_prefix_begin / pg_ar_loglik_prefix): per layer, the wild type's raw q/k/v rows (conv
  // look-back) and conv'd q/k/v rows (keys / values of the shared prefix), [layers][prefix_T][3d*np] fp16 each
  unsigned int* pack_scratch = nullptr;  // one word per packed matrix (absmax of the matrix), [layers * 6]
  int pack_used = 0;
  __half *raw_cache = nullptr, *kv_cache = nullptr;
  int prefix_T = 0;        // rows recorded by the last pg_ar_prefix_begin (0 = none)
  int prefix_cap = 0;      // rows the caches can hold
  // MSA Transformer: column-attention blocks (ln1*, wqkv, wo of a Layer), row-position table, tied-attention workspace
  std::vector<pg::Layer> col_layers;
  float* row_pos = nullptr;         // [1024, d] or null
  void* tied_ws = nullptr;          // one allocation, regrown on demand (not in `allocs`)
  size_t tied_bytes = 0;
  // compact buffers for the pruned last layer (one row per sequence)
  float* xc = nullptr;
  __half *cabuf = nullptr, *cfbuf = nullptr;
  // delta-operand mode (PG_PREC_F16D): per-layer base rows of the unmasked window (one allocation, regrown on demand, not in
  // `allocs`) and the compact exact q/k/v rows of the masked positions
  bool delta = false;
  float* base = nullptr;
  size_t base_floats = 0;  // bytes allocated for `base`
  int base_T = 0;
  __half* cq = nullptr;
};

namespace pg {
namespace {

int fail(pg_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return set_error(code, msg);
}

template <typename T>
int dev_alloc(pg_handle h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) return fail(h, PG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return PG_OK;
}

__global__ void row_select_kernel(const int32_t* positions, const int32_t* win_start, const int32_t* out_row, int p_offset, int n,
                                  int32_t* sel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int gp = p_offset + i;
  const int tok = out_row ? out_row[gp] : positions[gp];
  sel[i] = tok - (win_start ? win_start[gp] : 0);
}

// One linear layer of the model: A in the handle's operand format (common.h: fmt 0 / 1 / 2 by precision mode).
struct Lin {
  const __half* a; int64_t lda; float a_scale;  // a_scale: scale of A's e4m3 planes (PG_PREC_F16F8)
  const __half* w; const float* w_inv; const float* bias;
  int M, N, K, epi;
  __half* out = nullptr; int out_fmt = 0; float out_scale = 0.f;  // epi != 2
  float* resid = nullptr;                                          // epi == 2
};
int run_lin(pg_handle h, int cat, const Lin& L, cudaStream_t s, int rot_T = 0) {
  const int np = h->np;
  GemmLaunch g{};
  g.a = L.a; g.lda = L.lda; g.w = L.w; g.ldw = static_cast<int64_t>(L.K) * np; g.bias = L.bias;
  g.M = L.M; g.N = L.N; g.K = L.K; g.nseg = h->nseg; g.epi = L.epi;
  g.a_scale = L.a_scale; g.w_inv = L.w_inv; g.w_uniform = 1;  // pg_load_weights: one e4m3 scale per (fused) weight matrix
  if (L.epi == 2) {
    g.resid = L.resid; g.ldr = L.N;
  } else {
    g.out = L.out; g.ldo = static_cast<int64_t>(L.N) * np; g.out_fmt = L.out_fmt; g.out_lo_off = L.out_fmt ? L.N : 0;
    g.out_scale = L.out_scale;
  }
  if (L.epi == 3) { g.rot_cos = h->rot_cos; g.rot_sin = h->rot_sin; g.rot_T = rot_T; g.rot_dim = h->desc.embed_dim; }
  ProfScope ps(cat, s);
  return launch_gemm(g, s);
}

// `emit_rows` (device, [Bc]) = the one row per sequence the caller will read, or null when every row is needed. When given,
// the last layer runs its attention-output / out_proj / LayerNorm / MLP for those rows only and leaves them in h->xc [Bc, d].
int forward_rows(pg_handle h, const int32_t* tokens, int n_tokens, const int32_t* positions, const int32_t* win_start,
                 int p_offset, int Bc, int T, cudaStream_t s, const int32_t* emit_rows = nullptr) {
  const pg_model_desc& D = h->desc;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np;
  const int afmt = h->nseg == 2 ? 2 : (np == 2 ? 1 : 0);  // operand format of GEMM inputs
  const int qfmt = np == 2 ? 1 : 0;                        // q/k/v for the attention kernel: fp16 hi [| lo]
  const int64_t ldd = static_cast<int64_t>(d) * np, ldf = static_cast<int64_t>(f) * np, ldq = static_cast<int64_t>(3 * d) * np;
  const int rows = Bc * T;
  EmbedLaunch e{};
  e.tokens = tokens; e.n_tokens = n_tokens; e.positions = positions; e.win_start = win_start;
  e.P = Bc; e.T = T; e.d = d; e.embed = h->embed; e.pos_table = (D.arch == PG_ARCH_ESM1B) ? h->pos : nullptr;
  e.lnb_gamma = D.emb_ln_before ? h->lnbg : nullptr; e.lnb_beta = D.emb_ln_before ? h->lnbb : nullptr;
  e.token_dropout = D.token_dropout; e.mask_idx = 32; e.p_offset = p_offset; e.x = h->x;
  int rc;
  { ProfScope ps(CAT_EMBED, s); rc = launch_embed(e, s); }
  if (rc) return rc;
  const bool rotary = (D.arch == PG_ARCH_ESM2);
  auto ln = [&](const float* x, const float* g, const float* b, int nrows, __half* out) {
    ProfScope ps(CAT_LN, s);
    return launch_layernorm_f16(x, d, g, b, nrows, d, out, ldd, np == 2 ? d : 0, s, afmt, S_LN);
  };
  for (int l = 0; l < D.layers; ++l) {
    const Layer& L = h->layers[l];
    rc = ln(h->x, L.ln1g, L.ln1b, rows, h->abuf);
    if (rc) return rc;
    Lin q{h->abuf, ldd, S_LN, L.wqkv, L.iqkv, L.bqkv, rows, 3 * d, d, rotary ? 3 : 0};
    q.out = h->qkv; q.out_fmt = qfmt;
    rc = run_lin(h, CAT_GEMM_QKV, q, s, T);
    if (rc) return rc;
    if (emit_rows && l == D.layers - 1) {
      // exact pruning of the final layer: one query row per sequence from here on
      { ProfScope ps(CAT_ATTN, s, 2);
        rc = launch_attn_single_query(h->qkv, ldq, np == 2 ? 3 * d : 0, emit_rows, Bc, T, D.heads, h->cabuf, ldd, np == 2 ? d : 0, s,
                                      afmt, S_ATT);
        if (!rc) rc = launch_gather_rows(h->x, emit_rows, Bc, T, d, h->xc, s); }
      if (rc) return rc;
      Lin o{h->cabuf, ldd, S_ATT, L.wo, L.io, L.bo, Bc, d, d, 2};
      o.resid = h->xc;
      rc = run_lin(h, CAT_GEMM_OUT, o, s);
      if (rc) return rc;
      rc = ln(h->xc, L.ln2g, L.ln2b, Bc, h->cabuf);
      if (rc) return rc;
      Lin f1{h->cabuf, ldd, S_LN, L.w1, L.i1, L.b1, Bc, f, d, 1};
      f1.out = h->cfbuf; f1.out_fmt = afmt; f1.out_scale = S_GELU;
      rc = run_lin(h, CAT_GEMM_FC1, f1, s);
      if (rc) return rc;
      Lin f2{h->cfbuf, ldf, S_GELU, L.w2, L.i2, L.b2, Bc, d, f, 2};
      f2.resid = h->xc;
      return run_lin(h, CAT_GEMM_FC2, f2, s);
    }
    AttnLaunch a{};
    a.qkv = h->qkv; a.ld = ldq; a.lo_off = np == 2 ? 3 * d : 0;
    a.out = h->abuf; a.ldo = ldd; a.out_lo_off = np == 2 ? d : 0; a.out_fmt = afmt; a.out_scale = S_ATT;
    a.B = Bc; a.T = T; a.heads = D.heads; a.nseg = np == 2 ? 3 : 1; a.causal = 0; a.alibi_slopes = nullptr;
    { ProfScope ps(CAT_ATTN, s); rc = launch_attention_tc(a, s); }
    if (rc) return rc;
    Lin o{h->abuf, ldd, S_ATT, L.wo, L.io, L.bo, rows, d, d, 2};
    o.resid = h->x;
    rc = run_lin(h, CAT_GEMM_OUT, o, s);
    if (rc) return rc;
    rc = ln(h->x, L.ln2g, L.ln2b, rows, h->abuf);
    if (rc) return rc;
    Lin f1{h->abuf, ldd, S_LN, L.w1, L.i1, L.b1, rows, f, d, 1};
    f1.out = h->fbuf; f1.out_fmt = afmt; f1.out_scale = S_GELU;
    rc = run_lin(h, CAT_GEMM_FC1, f1, s);
    if (rc) return rc;
    Lin f2{h->fbuf, ldf, S_GELU, L.w2, L.i2, L.b2, rows, d, f, 2};
    f2.resid = h->x;
    rc = run_lin(h, CAT_GEMM_FC2, f2, s);
    if (rc) return rc;
  }
  return PG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Delta-operand mode (PG_PREC_F16D) of the masked-marginal pass. The P masked copies of a window differ from the unmasked window
// by a perturbation: every GEMM input is a = a0[t] + D with the base row a0[t] shared by all copies. The linear layers then are
//   W a + b = (W a0[t] + b)  +  W D
// with the first term computed ONCE per window at the fp16 hi/lo x3 precision (forward_base: T rows) and the second per copy in a
// SINGLE fp16 pass on D = rn16(a - a0[t]) — the products' error scales with |D|, which is 6-10 % of |a| on the rows that do not
// hold the mask (scripts/precision_delta.py, true ESM-1v 650M size: max score error 5.7e-4 at L = 96 against 1.5e-2 for a plain
// single pass). The row that holds the mask (|D| ~ |a|) takes the compact exact path: P rows per layer through the x3 GEMMs, like
// the pruned last layer. One tensor-pipe unit per algorithmic FLOP instead of two (fp16 + e4m3 cross terms) or three.
//   base rows per layer, T rows each. Inputs of the four GEMMs, as ONE fp16 plane (any fixed reference works as long as the matching
//   output rows were computed from exactly these values): a0 = rn16(LN1 output), o0 = rn16(attention output), b0 = rn16(LN2 output),
//   f0 = rn16(GELU output). Their images, fp32: bqkv = Wqkv a0 + b, bout = Wo o0 + b, bfc1 = W1 b0 + b (pre-GELU), bfc2 = W2 f0 + b.
struct BaseRows { const float *bqkv, *bout, *bfc1, *bfc2; const __half *a0, *o0, *b0, *f0; };
size_t base_bytes(const pg_model_desc& D, int T) {
  const size_t d = D.embed_dim, f = D.ffn_dim;
  return static_cast<size_t>(D.layers) * T * (4 * (5 * d + f) + 2 * (3 * d + f));
}
BaseRows base_rows(pg_handle h, int l) {
  const long long T = h->base_T, d = h->desc.embed_dim, f = h->desc.ffn_dim;
  const float* p = h->base + static_cast<long long>(l) * T * (5 * d + f);
  const __half* q = reinterpret_cast<const __half*>(h->base + static_cast<long long>(h->desc.layers) * T * (5 * d + f)) +
                    static_cast<long long>(l) * T * (3 * d + f);
  BaseRows b;
  b.bqkv = p; b.bout = p + T * 3 * d; b.bfc1 = p + T * 4 * d; b.bfc2 = p + T * (4 * d + f);
  b.a0 = q; b.o0 = q + T * d; b.b0 = q + T * 2 * d; b.f0 = q + T * 3 * d;
  return b;
}

// The unmasked window (with the token-dropout scale of a copy that holds one mask) through the model at x3 precision, recording
// the base rows of every layer. Uses the pass workspace (T rows of it) before the passes start.
int forward_base(pg_handle h, const int32_t* tokens, int n_tokens, int T, cudaStream_t s) {
  const pg_model_desc& D = h->desc;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np;
  const int64_t ldd = static_cast<int64_t>(d) * np, ldf = static_cast<int64_t>(f) * np, ldq = static_cast<int64_t>(3 * d) * np;
  const size_t need = base_bytes(D, T);
  if (need > h->base_floats) {  // (base_floats counts bytes)
    if (h->base) cudaFree(h->base);
    h->base = nullptr; h->base_floats = 0;
    cudaError_t e = cudaMalloc(&h->base, need);
    if (e != cudaSuccess) return set_error(PG_ERR_CUDA, std::string("forward_base: cudaMalloc: ") + cudaGetErrorString(e));
    h->base_floats = need;
  }
  h->base_T = T;
  PG_CUDA_OK(cudaMemsetAsync(h->base, 0, need, s));  // the tap GEMMs reduce-add into zero
  EmbedLaunch e{};
  e.tokens = tokens; e.n_tokens = n_tokens; e.positions = nullptr; e.win_start = nullptr;
  e.P = 1; e.T = T; e.d = d; e.embed = h->embed; e.pos_table = h->pos;
  e.lnb_gamma = D.emb_ln_before ? h->lnbg : nullptr; e.lnb_beta = D.emb_ln_before ? h->lnbb : nullptr;
  e.token_dropout = D.token_dropout; e.mask_idx = 32; e.p_offset = 0; e.x = h->x; e.force_mask_scale = 1;
  int rc;
  { ProfScope ps(CAT_EMBED, s); rc = launch_embed(e, s); }
  if (rc) return rc;
  auto ln = [&](const float* g, const float* b) {
    ProfScope ps(CAT_LN, s);
    return launch_layernorm_f16(h->x, d, g, b, T, d, h->abuf, ldd, d, s, 1, 0.f);
  };
  auto tap = [&](int cat, const __half* a, int64_t lda, const __half* w, const float* bias, int N, int K, const float* dst) {
    Lin t{a, lda, 0.f, w, nullptr, bias, T, N, K, 2};
    t.resid = const_cast<float*>(dst);
    return run_lin(h, cat, t, s);
  };
  // The base row of a GEMM input is its fp16 hi plane: zero the lo plane of the T rows (the tap GEMM then multiplies exactly that
  // plane by the full hi/lo weights) and keep a copy of the hi plane.
  auto fix_base = [&](__half* buf, int64_t ld, int n, const __half* dst) -> int {
    ProfScope ps(CAT_OTHER, s);
    PG_CUDA_OK(cudaMemset2DAsync(buf + n, ld * 2, 0, static_cast<size_t>(n) * 2, T, s));
    PG_CUDA_OK(cudaMemcpy2DAsync(const_cast<__half*>(dst), static_cast<size_t>(n) * 2, buf, ld * 2, static_cast<size_t>(n) * 2, T,
                                 cudaMemcpyDeviceToDevice, s));
    return PG_OK;
  };
  for (int l = 0; l < D.layers; ++l) {
    const Layer& L = h->layers[l];
    const BaseRows B = base_rows(h, l);
    if ((rc = ln(L.ln1g, L.ln1b))) return rc;
    Lin q{h->abuf, ldd, 0.f, L.wqkv, nullptr, L.bqkv, T, 3 * d, d, 0};
    q.out = h->qkv; q.out_fmt = 1;
    if ((rc = run_lin(h, CAT_GEMM_QKV, q, s))) return rc;
    if ((rc = fix_base(h->abuf, ldd, d, B.a0))) return rc;
    if ((rc = tap(CAT_GEMM_QKV, h->abuf, ldd, L.wqkv, L.bqkv, 3 * d, d, B.bqkv))) return rc;
    AttnLaunch a{};
    a.qkv = h->qkv; a.ld = ldq; a.lo_off = 3 * d;
    a.out = h->abuf; a.ldo = ldd; a.out_lo_off = d; a.out_fmt = 1;
    a.B = 1; a.T = T; a.heads = D.heads; a.nseg = 3; a.causal = 0; a.alibi_slopes = nullptr;
    { ProfScope ps(CAT_ATTN, s); rc = launch_attention_tc(a, s); }
    if (rc) return rc;
    Lin o{h->abuf, ldd, 0.f, L.wo, nullptr, L.bo, T, d, d, 2};
    o.resid = h->x;
    if ((rc = run_lin(h, CAT_GEMM_OUT, o, s))) return rc;
    if ((rc = fix_base(h->abuf, ldd, d, B.o0))) return rc;
    if ((rc = tap(CAT_GEMM_OUT, h->abuf, ldd, L.wo, L.bo, d, d, B.bout))) return rc;
    if ((rc = ln(L.ln2g, L.ln2b))) return rc;
    Lin f1{h->abuf, ldd, 0.f, L.w1, nullptr, L.b1, T, f, d, 1};
    f1.out = h->fbuf; f1.out_fmt = 1;
    if ((rc = run_lin(h, CAT_GEMM_FC1, f1, s))) return rc;
    if ((rc = fix_base(h->abuf, ldd, d, B.b0))) return rc;
    if ((rc = tap(CAT_GEMM_FC1, h->abuf, ldd, L.w1, L.b1, f, d, B.bfc1))) return rc;
    Lin f2{h->fbuf, ldf, 0.f, L.w2, nullptr, L.b2, T, d, f, 2};
    f2.resid = h->x;
    if ((rc = run_lin(h, CAT_GEMM_FC2, f2, s))) return rc;
    if ((rc = fix_base(h->fbuf, ldf, f, B.f0))) return rc;
    if ((rc = tap(CAT_GEMM_FC2, h->fbuf, ldf, L.w2, L.b2, d, f, B.bfc2))) return rc;
  }
  return PG_OK;
}

// One single-pass GEMM on difference rows: C = epi(base_pre[t] + A_delta * W_hi^T) [- base_post[t]].
int run_lin_delta(pg_handle h, int cat, const __half* a, int64_t lda, const __half* w, int M, int N, int K, int epi, const float* base_pre,
                  const __half* base_post, int T, const int32_t* mask_pos, __half* out, int64_t ldo, int out_fmt, float* resid,
                  cudaStream_t s) {
  GemmLaunch g{};
  g.a = a; g.lda = lda; g.w = w; g.ldw = static_cast<int64_t>(K) * h->np; g.bias = nullptr;  // W: the hi plane of the hi/lo rows
  g.M = M; g.N = N; g.K = K; g.nseg = 1; g.epi = epi;
  g.base_pre = base_pre; g.base_post = base_post; g.base_T = T; g.mask_pos = mask_pos;
  if (epi == 2) {
    g.resid = resid; g.ldr = N;
  } else {
    g.out = out; g.ldo = ldo; g.out_fmt = out_fmt; g.out_lo_off = out_fmt ? N : 0;
  }
  ProfScope ps(cat, s);
  return launch_gemm(g, s);
}

// forward_rows in delta-operand mode (single window: every copy sees tokens [0, T); emit_rows[b] = the masked token of copy b).
int forward_rows_delta(pg_handle h, const int32_t* tokens, int n_tokens, const int32_t* positions, const int32_t* win_start,
                       int p_offset, int Bc, int T, cudaStream_t s, const int32_t* emit_rows) {
  const pg_model_desc& D = h->desc;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np;
  const int64_t ldd = static_cast<int64_t>(d) * np, ldf = static_cast<int64_t>(f) * np, ldq = static_cast<int64_t>(3 * d) * np;
  const int rows = Bc * T;
  EmbedLaunch e{};
  e.tokens = tokens; e.n_tokens = n_tokens; e.positions = positions; e.win_start = win_start;
  e.P = Bc; e.T = T; e.d = d; e.embed = h->embed; e.pos_table = h->pos;
  e.lnb_gamma = D.emb_ln_before ? h->lnbg : nullptr; e.lnb_beta = D.emb_ln_before ? h->lnbb : nullptr;
  e.token_dropout = D.token_dropout; e.mask_idx = 32; e.p_offset = p_offset; e.x = h->x;
  int rc;
  { ProfScope ps(CAT_EMBED, s); rc = launch_embed(e, s); }
  if (rc) return rc;
  { ProfScope ps(CAT_OTHER, s); rc = launch_gather_rows(h->x, emit_rows, Bc, T, d, h->xc, s); }  // xc = the masked rows, kept exact
  if (rc) return rc;
  auto ln_delta = [&](const float* g, const float* b, const __half* base) {   // all rows: LN(x) - base[t] -> abuf (fp16, pitch ldd)
    ProfScope ps(CAT_LN, s);
    return launch_layernorm_f16(h->x, d, g, b, rows, d, h->abuf, ldd, 0, s, 0, 0.f, 0, 0, base, T);
  };
  auto ln_compact = [&](const float* g, const float* b) {                     // masked rows: LN(xc) -> cabuf (hi | lo)
    ProfScope ps(CAT_LN, s);
    return launch_layernorm_f16(h->xc, d, g, b, Bc, d, h->cabuf, ldd, d, s, 1, 0.f);
  };
  auto put_back = [&]() {                                                    // x[masked rows] <- xc
    ProfScope ps(CAT_OTHER, s);
    return launch_scatter_rows(h->xc, static_cast<int64_t>(d) * 4, h->x, static_cast<int64_t>(d) * 4, emit_rows, Bc, T, d * 4, s);
  };
  for (int l = 0; l < D.layers; ++l) {
    const Layer& L = h->layers[l];
    const BaseRows B = base_rows(h, l);
    // q/k/v: all rows from the difference, the masked rows exactly
    if ((rc = ln_delta(L.ln1g, L.ln1b, B.a0))) return rc;
    if ((rc = run_lin_delta(h, CAT_GEMM_QKV, h->abuf, ldd, L.wqkv, rows, 3 * d, d, 0, B.bqkv, nullptr, T, nullptr, h->qkv, ldq, 1, nullptr, s)))
      return rc;
    if ((rc = ln_compact(L.ln1g, L.ln1b))) return rc;
    Lin cq{h->cabuf, ldd, 0.f, L.wqkv, nullptr, L.bqkv, Bc, 3 * d, d, 0};
    cq.out = h->cq; cq.out_fmt = 1;
    if ((rc = run_lin(h, CAT_GEMM_QKV, cq, s))) return rc;
    { ProfScope ps(CAT_OTHER, s);
      rc = launch_scatter_rows(h->cq, ldq * 2, h->qkv, ldq * 2, emit_rows, Bc, T, static_cast<int>(ldq * 2), s); }
    if (rc) return rc;
    if (l == D.layers - 1) {
      // exact pruning of the final layer, as in forward_rows: one query row per copy from here on, x3 GEMMs on full values
      { ProfScope ps(CAT_ATTN, s, 2);
        rc = launch_attn_single_query(h->qkv, ldq, 3 * d, emit_rows, Bc, T, D.heads, h->cabuf, ldd, d, s, 1, 0.f); }
      if (rc) return rc;
      Lin o{h->cabuf, ldd, 0.f, L.wo, nullptr, L.bo, Bc, d, d, 2};
      o.resid = h->xc;
      if ((rc = run_lin(h, CAT_GEMM_OUT, o, s))) return rc;
      if ((rc = ln_compact(L.ln2g, L.ln2b))) return rc;
      Lin f1{h->cabuf, ldd, 0.f, L.w1, nullptr, L.b1, Bc, f, d, 1};
      f1.out = h->cfbuf; f1.out_fmt = 1;
      if ((rc = run_lin(h, CAT_GEMM_FC1, f1, s))) return rc;
      Lin f2{h->cfbuf, ldf, 0.f, L.w2, nullptr, L.b2, Bc, d, f, 2};
      f2.resid = h->xc;
      return run_lin(h, CAT_GEMM_FC2, f2, s);
    }
    // attention on full-value q/k/v; output as difference rows, the masked rows' exact output to cabuf
    AttnLaunch a{};
    a.qkv = h->qkv; a.ld = ldq; a.lo_off = 3 * d;
    a.out = h->abuf; a.ldo = ldd; a.out_lo_off = 0; a.out_fmt = 0;
    a.B = Bc; a.T = T; a.heads = D.heads; a.nseg = 3; a.causal = 0; a.alibi_slopes = nullptr;
    a.base_o = B.o0; a.mask_pos = emit_rows; a.cout = h->cabuf; a.ldc = ldd; a.c_lo_off = d;
    { ProfScope ps(CAT_ATTN, s); rc = launch_attention_tc(a, s); }
    if (rc) return rc;
    if ((rc = run_lin_delta(h, CAT_GEMM_OUT, h->abuf, ldd, L.wo, rows, d, d, 2, B.bout, nullptr, T, emit_rows, nullptr, 0, 0, h->x, s))) return rc;
    Lin o{h->cabuf, ldd, 0.f, L.wo, nullptr, L.bo, Bc, d, d, 2};
    o.resid = h->xc;
    if ((rc = run_lin(h, CAT_GEMM_OUT, o, s))) return rc;
    if ((rc = put_back())) return rc;
    // MLP
    if ((rc = ln_delta(L.ln2g, L.ln2b, B.b0))) return rc;
    if ((rc = run_lin_delta(h, CAT_GEMM_FC1, h->abuf, ldd, L.w1, rows, f, d, 1, B.bfc1, B.f0, T, nullptr, h->fbuf, ldf, 0, nullptr, s))) return rc;
    if ((rc = ln_compact(L.ln2g, L.ln2b))) return rc;
    Lin f1{h->cabuf, ldd, 0.f, L.w1, nullptr, L.b1, Bc, f, d, 1};
    f1.out = h->cfbuf; f1.out_fmt = 1;
    if ((rc = run_lin(h, CAT_GEMM_FC1, f1, s))) return rc;
    if ((rc = run_lin_delta(h, CAT_GEMM_FC2, h->fbuf, ldf, L.w2, rows, d, f, 2, B.bfc2, nullptr, T, emit_rows, nullptr, 0, 0, h->x, s))) return rc;
    Lin f2{h->cfbuf, ldf, 0.f, L.w2, nullptr, L.b2, Bc, d, f, 2};
    f2.resid = h->xc;
    if ((rc = run_lin(h, CAT_GEMM_FC2, f2, s))) return rc;
    if ((rc = put_back())) return rc;
  }
  return PG_OK;
}

// Tranception decoder stack over B right-padded sequences of T tokens (model_pytorch.py:526-612): no padding mask is needed
// because pads sit to the right of every real token and attention is causal.
// mode 0: plain forward. mode 1 (B == 1): also record every layer's raw and conv'd q/k/v rows into the prefix caches.
// mode 2: the B sequences are rows [start, start + T) of sequences whose first `start` rows (multiple of 128) equal the recorded
// wild type's: depthwise-conv look-back and the keys / values of the prefix come from the caches.
int forward_tranception(pg_handle h, const int32_t* ids, int B, int T, cudaStream_t s, int mode = 0, int start = 0) {
  const pg_model_desc& D = h->desc;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np;
  const int afmt = h->nseg == 2 ? 2 : (np == 2 ? 1 : 0);
  const int qfmt = np == 2 ? 1 : 0;
  const int64_t ldd = static_cast<int64_t>(d) * np, ldf = static_cast<int64_t>(f) * np, ldq = static_cast<int64_t>(3 * d) * np;
  const int rows = B * T;
  int rc;
  { ProfScope ps(CAT_EMBED, s); rc = launch_gather_embed(ids, h->embed, rows, d, D.vocab, h->x, s); }
  if (rc) return rc;
  auto ln = [&](const float* g, const float* b) {
    ProfScope ps(CAT_LN, s);
    return launch_layernorm_f16(h->x, d, g, b, rows, d, h->abuf, ldd, np == 2 ? d : 0, s, afmt, S_LN);
  };
  for (int l = 0; l < D.layers; ++l) {
    const Layer& L = h->layers[l];
    rc = ln(L.ln1g, L.ln1b);
    if (rc) return rc;
    Lin q{h->abuf, ldd, S_LN, L.wqkv, L.iqkv, L.bqkv, rows, 3 * d, d, 0};
    q.out = h->qkv; q.out_fmt = qfmt;
    rc = run_lin(h, CAT_GEMM_QKV, q, s);
    if (rc) return rc;
    const size_t layer_off = static_cast<size_t>(l) * h->prefix_cap * ldq;
    if (mode == 1)
      PG_CUDA_OK(cudaMemcpyAsync(h->raw_cache + layer_off, h->qkv, static_cast<size_t>(T) * ldq * sizeof(__half), cudaMemcpyDeviceToDevice, s));
    { ProfScope ps(CAT_OTHER, s);
      rc = launch_qkv_conv(h->qkv, h->qkv2, ldq, np == 2 ? 3 * d : 0, B, T, D.heads, L.conv_taps, 0.125f, s,
                           mode == 2 ? h->raw_cache + layer_off + static_cast<size_t>(start - 6) * ldq : nullptr); }
    if (rc) return rc;
    if (mode == 1)
      PG_CUDA_OK(cudaMemcpyAsync(h->kv_cache + layer_off, h->qkv2, static_cast<size_t>(T) * ldq * sizeof(__half), cudaMemcpyDeviceToDevice, s));
    AttnLaunch a{};
    a.qkv = h->qkv2; a.ld = ldq; a.lo_off = np == 2 ? 3 * d : 0;
    a.out = h->abuf; a.ldo = ldd; a.out_lo_off = np == 2 ? d : 0; a.out_fmt = afmt; a.out_scale = S_ATT;
    a.B = B; a.T = T; a.heads = D.heads; a.nseg = np == 2 ? 3 : 1; a.causal = 1; a.alibi_slopes = h->slopes;
    if (mode == 2) { a.prefix = h->kv_cache + layer_off; a.prefix_len = start; }
    { ProfScope ps(CAT_ATTN, s); rc = launch_attention_tc(a, s); }
    if (rc) return rc;
    Lin o{h->abuf, ldd, S_ATT, L.wo, L.io, L.bo, rows, d, d, 2};
    o.resid = h->x;
    rc = run_lin(h, CAT_GEMM_OUT, o, s);
    if (rc) return rc;
    rc = ln(L.ln2g, L.ln2b);
    if (rc) return rc;
    Lin f1{h->abuf, ldd, S_LN, L.w1, L.i1, L.b1, rows, f, d, 4};
    f1.out = h->fbuf; f1.out_fmt = afmt; f1.out_scale = S_RELU2;
    rc = run_lin(h, CAT_GEMM_FC1, f1, s);
    if (rc) return rc;
    Lin f2{h->fbuf, ldf, S_RELU2, L.w2, L.i2, L.b2, rows, d, f, 2};
    f2.resid = h->x;
    rc = run_lin(h, CAT_GEMM_FC2, f2, s);
    if (rc) return rc;
  }
  return PG_OK;
}

// MSA Transformer stack (msa_transformer.py:150-222) over Bc masked copies of an [R, Cw] alignment window; token rows are (b, r, c)
// with c fastest. Per layer (AxialTransformerLayer, modules.py:205-235): tied row attention, column attention, FFN, each a pre-LN
// residual block. The tied attention's two products run as grouped GEMMs on regrouped operands (msa_transformer.cu), always with fp16
// hi/lo pairs when the handle has them (x3: they are < 10 % of the FLOPs); the column attention is the tcgen05 attention kernel over
// the B*Cw columns as sequences of R rows, fed by a LayerNorm that writes its rows in (b, c, r) order.
// `sel` (device, [Bc]): the window column of each alignment whose (row 0) output the caller reads. The last layer then runs exactly
// what that output depends on — the full q/k/v projection of its row attention, one row of each tied attention map, and the rest on
// the Bc*R rows of column sel[b] (column attention) and on Bc rows (FFN); the result is left in h->xc [Bc, d].
int forward_msa(pg_handle h, const int32_t* tokens, int R, int Cfull, const int32_t* positions, const int32_t* win_start, int p_offset,
                int Bc, int Cw, cudaStream_t s, const int32_t* sel) {
  const pg_model_desc& D = h->desc;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np, H = D.heads;
  const int afmt = h->nseg == 2 ? 2 : (np == 2 ? 1 : 0);
  const int qfmt = np == 2 ? 1 : 0;
  const int64_t ldd = static_cast<int64_t>(d) * np, ldf = static_cast<int64_t>(f) * np, ldq = static_cast<int64_t>(3 * d) * np;
  const int rows = Bc * R * Cw;
  // tied-attention workspace: Q' / K' [G*Cp, Nt*np] (Q' doubles as the context buffer), V' [G*Nt, Kp*np], S fp32 [G*Cp, Kp], P [G*Cp, Kp*np]
  const int G = Bc * H, Cp = (Cw + 127) / 128 * 128, Kp = (Cw + 63) / 64 * 64, Nt = R * 64;
  const int64_t ldt = static_cast<int64_t>(Nt) * np, ldv = static_cast<int64_t>(Kp) * np, lds = Kp, ldp = static_cast<int64_t>(Kp) * np;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t b_q = al(static_cast<size_t>(G) * Cp * ldt * 2), b_v = al(static_cast<size_t>(G) * Nt * ldv * 2);
  const size_t b_s = al(static_cast<size_t>(G) * Cp * lds * 4), b_p = al(static_cast<size_t>(G) * Cp * ldp * 2);
  const size_t need = 2 * b_q + b_v + b_s + b_p;
  if (need > h->tied_bytes) {
    if (h->tied_ws) cudaFree(h->tied_ws);
    h->tied_ws = nullptr; h->tied_bytes = 0;
    PG_CUDA_OK(cudaMalloc(&h->tied_ws, need));
    h->tied_bytes = need;
    PG_CUDA_OK(cudaMemsetAsync(h->tied_ws, 0, need, s));
  }
  uint8_t* ws = static_cast<uint8_t*>(h->tied_ws);
  __half* tq = reinterpret_cast<__half*>(ws);
  __half* tk = reinterpret_cast<__half*>(ws + b_q);
  __half* tv = reinterpret_cast<__half*>(ws + 2 * b_q);
  float* S = reinterpret_cast<float*>(ws + 2 * b_q + b_v);
  __half* P = reinterpret_cast<__half*>(ws + 2 * b_q + b_v + b_s);
  int rc;
  MsaEmbedLaunch e{};
  e.tokens = tokens; e.R = R; e.Cfull = Cfull; e.positions = positions; e.win_start = win_start; e.p_offset = p_offset;
  e.B = Bc; e.Cw = Cw; e.d = d; e.embed = h->embed; e.pos_table = h->pos; e.row_pos = h->row_pos;
  e.gamma = h->lnbg; e.beta = h->lnbb; e.mask_idx = 32; e.x = h->x;
  { ProfScope ps(CAT_EMBED, s); rc = launch_msa_embed(e, s); }
  if (rc) return rc;
  auto ln = [&](const float* g, const float* b, int pR, int pC) {
    ProfScope ps(CAT_LN, s);
    return launch_layernorm_f16(h->x, d, g, b, rows, d, h->abuf, ldd, np == 2 ? d : 0, s, afmt, S_LN, pR, pC);
  };
  const float row_scale = 1.0f / sqrtf(static_cast<float>(R));  // align_scaling (axial_attention.py:78-80); head_dim^-1/2 is in the weights
  for (int l = 0; l < D.layers; ++l) {
    const Layer& L = h->layers[l];
    const Layer& CL = h->col_layers[l];
    // ---- tied row attention
    rc = ln(L.ln1g, L.ln1b, 0, 0);
    if (rc) return rc;
    Lin q{h->abuf, ldd, S_LN, L.wqkv, L.iqkv, L.bqkv, rows, 3 * d, d, 0};
    q.out = h->qkv; q.out_fmt = qfmt;
    rc = run_lin(h, CAT_GEMM_QKV, q, s);
    if (rc) return rc;
    if (l == D.layers - 1) {
      // ---- pruned last layer (exact)
      const int BR = Bc * R;
      uint8_t* fb = reinterpret_cast<uint8_t*>(h->fbuf);  // the MLP hidden buffer is idle here: compact residual / operand / q-k-v rows
      float* xr = reinterpret_cast<float*>(fb);
      __half* cab = reinterpret_cast<__half*>(fb + al(static_cast<size_t>(BR) * d * 4));
      __half* cqkv = reinterpret_cast<__half*>(fb + al(static_cast<size_t>(BR) * d * 4) + al(static_cast<size_t>(BR) * ldd * 2));
      { ProfScope ps(CAT_TIED, s, 3);
        rc = launch_tied_col_attention(h->qkv, ldq, np == 2 ? 3 * d : 0, sel, Bc, R, Cw, H, row_scale, S, cab, ldd, np == 2 ? d : 0, afmt,
                                       S_ATT, s); }
      if (rc) return rc;
      { ProfScope ps(CAT_REGROUP, s);
        rc = launch_msa_gather_col(h->x, sel, Bc, R, Cw, d, xr, s); }
      if (rc) return rc;
      Lin o{cab, ldd, S_ATT, L.wo, L.io, L.bo, BR, d, d, 2};
      o.resid = xr;
      rc = run_lin(h, CAT_GEMM_OUT, o, s);
      if (rc) return rc;
      { ProfScope ps(CAT_LN, s);
        rc = launch_layernorm_f16(xr, d, CL.ln1g, CL.ln1b, BR, d, cab, ldd, np == 2 ? d : 0, s, afmt, S_LN); }
      if (rc) return rc;
      Lin q2{cab, ldd, S_LN, CL.wqkv, CL.iqkv, CL.bqkv, BR, 3 * d, d, 0};
      q2.out = cqkv; q2.out_fmt = qfmt;
      rc = run_lin(h, CAT_GEMM_QKV, q2, s);
      if (rc) return rc;
      AttnLaunch a{};
      a.qkv = cqkv; a.ld = ldq; a.lo_off = np == 2 ? 3 * d : 0;
      a.out = cab; a.ldo = ldd; a.out_lo_off = np == 2 ? d : 0; a.out_fmt = afmt; a.out_scale = S_ATT;
      a.B = Bc; a.T = R; a.heads = H; a.nseg = np == 2 ? 3 : 1; a.causal = 0; a.alibi_slopes = nullptr;
      { ProfScope ps(CAT_ATTN, s); rc = launch_attention_tc(a, s); }
      if (rc) return rc;
      Lin o2{cab, ldd, S_ATT, CL.wo, CL.io, CL.bo, BR, d, d, 2};
      o2.resid = xr;
      rc = run_lin(h, CAT_GEMM_OUT, o2, s);
      if (rc) return rc;
      rc = launch_gather_rows(xr, h->row_sel + h->head_cap, Bc, R, d, h->xc, s);  // row 0 of every alignment
      if (rc) return rc;
      { ProfScope ps(CAT_LN, s);
        rc = launch_layernorm_f16(h->xc, d, L.ln2g, L.ln2b, Bc, d, h->cabuf, ldd, np == 2 ? d : 0, s, afmt, S_LN); }
      if (rc) return rc;
      Lin f1{h->cabuf, ldd, S_LN, L.w1, L.i1, L.b1, Bc, f, d, 1};
      f1.out = h->cfbuf; f1.out_fmt = afmt; f1.out_scale = S_GELU;
      rc = run_lin(h, CAT_GEMM_FC1, f1, s);
      if (rc) return rc;
      Lin f2{h->cfbuf, ldf, S_GELU, L.w2, L.i2, L.b2, Bc, d, f, 2};
      f2.resid = h->xc;
      return run_lin(h, CAT_GEMM_FC2, f2, s);
    }
    { ProfScope ps(CAT_REGROUP, s, 2);
      rc = launch_tied_gather_qk(h->qkv, ldq, np == 2 ? 3 * d : 0, Bc, R, Cw, H, Cp, tq, tk, ldt, s);
      if (!rc) rc = launch_tied_transpose_v(h->qkv, ldq, np == 2 ? 3 * d : 0, Bc, R, Cw, H, Kp, tv, ldv, s); }
    if (rc) return rc;
    PG_CUDA_OK(cudaMemsetAsync(S, 0, b_s, s));
    { GemmLaunch g{};
      g.a = tq; g.lda = ldt; g.w = tk; g.ldw = ldt; g.M = G * Cp; g.N = Cw; g.K = Nt; g.nseg = np == 2 ? 3 : 1; g.epi = 2;
      g.resid = S; g.ldr = lds; g.grp_rows_a = Cp; g.grp_rows_b = Cp;
      ProfScope ps(CAT_TIED, s);
      rc = launch_gemm(g, s); }
    if (rc) return rc;
    { ProfScope ps(CAT_REGROUP, s);
      rc = launch_tied_softmax(S, lds, G, Cw, Cp, Kp, row_scale, P, ldp, np, s); }
    if (rc) return rc;
    { GemmLaunch g{};
      g.a = P; g.lda = ldp; g.w = tv; g.ldw = ldv; g.M = G * Cp; g.N = Nt; g.K = Kp; g.nseg = np == 2 ? 3 : 1; g.epi = 0;
      g.out = tq; g.ldo = ldt; g.out_fmt = afmt; g.out_lo_off = afmt ? Nt : 0; g.out_scale = S_ATT;
      g.grp_rows_a = Cp; g.grp_rows_b = Nt;
      ProfScope ps(CAT_TIED, s);
      rc = launch_gemm(g, s); }
    if (rc) return rc;
    { ProfScope ps(CAT_REGROUP, s);
      rc = launch_tied_scatter_out(tq, ldt, h->abuf, ldd, afmt, Bc, R, Cw, H, Cp, s); }
    if (rc) return rc;
    Lin o{h->abuf, ldd, S_ATT, L.wo, L.io, L.bo, rows, d, d, 2};
    o.resid = h->x;
    rc = run_lin(h, CAT_GEMM_OUT, o, s);
    if (rc) return rc;
    // ---- column attention: rows regrouped to (b, c, r) by the LayerNorm, back to (b, r, c) by the attention kernel
    rc = ln(CL.ln1g, CL.ln1b, R, Cw);
    if (rc) return rc;
    Lin q2{h->abuf, ldd, S_LN, CL.wqkv, CL.iqkv, CL.bqkv, rows, 3 * d, d, 0};
    q2.out = h->qkv; q2.out_fmt = qfmt;
    rc = run_lin(h, CAT_GEMM_QKV, q2, s);
    if (rc) return rc;
    AttnLaunch a{};
    a.qkv = h->qkv; a.ld = ldq; a.lo_off = np == 2 ? 3 * d : 0;
    a.out = h->abuf; a.ldo = ldd; a.out_lo_off = np == 2 ? d : 0; a.out_fmt = afmt; a.out_scale = S_ATT;
    a.B = Bc * Cw; a.T = R; a.heads = H; a.nseg = np == 2 ? 3 : 1; a.causal = 0; a.alibi_slopes = nullptr; a.perm_C = Cw;
    { ProfScope ps(CAT_ATTN, s); rc = launch_attention_tc(a, s); }
    if (rc) return rc;
    Lin o2{h->abuf, ldd, S_ATT, CL.wo, CL.io, CL.bo, rows, d, d, 2};
    o2.resid = h->x;
    rc = run_lin(h, CAT_GEMM_OUT, o2, s);
    if (rc) return rc;
    // ---- feed-forward
    rc = ln(L.ln2g, L.ln2b, 0, 0);
    if (rc) return rc;
    Lin f1{h->abuf, ldd, S_LN, L.w1, L.i1, L.b1, rows, f, d, 1};
    f1.out = h->fbuf; f1.out_fmt = afmt; f1.out_scale = S_GELU;
    rc = run_lin(h, CAT_GEMM_FC1, f1, s);
    if (rc) return rc;
    Lin f2{h->fbuf, ldf, S_GELU, L.w2, L.i2, L.b2, rows, d, f, 2};
    f2.resid = h->x;
    rc = run_lin(h, CAT_GEMM_FC2, f2, s);
    if (rc) return rc;
  }
  return PG_OK;
}

HeadLaunch head_args(pg_handle h, int T) {
  HeadLaunch hl{};
  hl.x = h->x; hl.d = h->desc.embed_dim; hl.T = T;
  hl.lna_g = h->lnag; hl.lna_b = h->lnab; hl.dense_w = h->hdw; hl.dense_b = h->hdb;
  hl.ln_g = h->hlng; hl.ln_b = h->hlnb; hl.out_w = h->embed; hl.out_b = h->hbias; hl.vocab = h->desc.vocab;
  hl.scratch_a = h->hs_a; hl.scratch_b = h->hs_b;
  return hl;
}

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" {

int pg_abi_version(void) { return 5; }

long long pg_launch_count(void) {
  std::lock_guard<std::mutex> lk(prof_mu());
  long long t = 0;
  for (int i = 0; i < CAT_COUNT; ++i) t += prof().launches[i];
  return t;
}

int pg_profile_begin(void) {
  std::lock_guard<std::mutex> lk(prof_mu());
  prof().on = true;
  prof().used = 0;
  return PG_OK;
}

// Stops profiling and returns, per category, the summed device time (ms) and number of timed scopes.
// The caller must have synchronised the stream(s) first.
int pg_profile_end(float* ms, int32_t* counts, int32_t ncat) {
  std::lock_guard<std::mutex> lk(prof_mu());
  ProfState& p = prof();
  p.on = false;
  for (int i = 0; i < ncat; ++i) { ms[i] = 0.f; counts[i] = 0; }
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]) != cudaSuccess) continue;
    const int c = p.ev_cat[i / 2];
    if (c < ncat) { ms[c] += t; counts[c] += 1; }
  }
  p.used = 0;
  return PG_OK;
}

const char* pg_last_error(pg_handle h) {
  if (h && !h->err.empty()) return h->err.c_str();
  return tls_error().c_str();
}

int pg_create(const pg_model_desc* desc, pg_handle* out) {
  if (!desc || !out) return set_error(PG_ERR_ARG, "pg_create: null argument");
  const pg_model_desc& D = *desc;
  if (D.layers <= 0 || D.embed_dim <= 0 || D.heads <= 0 || D.ffn_dim <= 0 || D.vocab <= 0)
    return set_error(PG_ERR_ARG, "pg_create: non-positive model dimension");
  if (D.embed_dim != D.heads * 64) return set_error(PG_ERR_UNSUPPORTED, "pg_create: head_dim must be 64");
  if (D.embed_dim % 64 || D.ffn_dim % 64) return set_error(PG_ERR_UNSUPPORTED, "pg_create: embed_dim and ffn_dim must be multiples of 64");
  if (D.arch != PG_ARCH_ESM1B && D.arch != PG_ARCH_ESM2 && D.arch != PG_ARCH_TRANCEPTION && D.arch != PG_ARCH_MSA)
    return set_error(PG_ERR_UNSUPPORTED, "pg_create: unknown arch");
  if (D.arch == PG_ARCH_TRANCEPTION && D.heads % 4) return set_error(PG_ERR_UNSUPPORTED, "pg_create: Tranception needs heads % 4 == 0 (model_pytorch.py:129-131)");
  if (D.precision != PG_PREC_F16 && D.precision != PG_PREC_F16X3 && D.precision != PG_PREC_F16F8 && D.precision != PG_PREC_F16D)
    return set_error(PG_ERR_ARG, "pg_create: unknown precision");
  if (D.precision == PG_PREC_F16D && D.arch != PG_ARCH_ESM1B)
    return set_error(PG_ERR_UNSUPPORTED, "pg_create: PG_PREC_F16D (delta operands) is built for the ESM-1b / ESM-1v masked-marginal path");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_error(PG_ERR_CUDA, "pg_create: no CUDA device (the B200 path has no CPU fallback)");
  if (D.device < 0 || D.device >= ndev) return set_error(PG_ERR_ARG, "pg_create: bad device ordinal");
  PG_CUDA_OK(cudaSetDevice(D.device));
  cudaDeviceProp prop;
  PG_CUDA_OK(cudaGetDeviceProperties(&prop, D.device));
  if (prop.major != 10) return set_error(PG_ERR_UNSUPPORTED, "pg_create: device is not sm_100 (tcgen05/TMEM required)");
  pg_handle h = new pg_handle_s();
  h->desc = D;
  h->np = (D.precision == PG_PREC_F16) ? 1 : 2;
  h->nseg = (D.precision == PG_PREC_F16) ? 1 : (D.precision == PG_PREC_F16F8 ? 2 : 3);  // F16D: the fp16 hi/lo x3 plumbing + delta GEMMs
  h->delta = (D.precision == PG_PREC_F16D);
  h->max_rows = D.max_rows > 0 ? D.max_rows : 131072;
  h->head_cap = 8192;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np;
  int rc = PG_OK;
  auto A = [&](auto** p, size_t n) { if (!rc) rc = dev_alloc(h, p, n); };
  h->layers.resize(D.layers);
  for (auto& L : h->layers) {
    A(&L.wqkv, static_cast<size_t>(3) * d * d * np); A(&L.wo, static_cast<size_t>(d) * d * np);
    A(&L.w1, static_cast<size_t>(f) * d * np); A(&L.w2, static_cast<size_t>(d) * f * np);
    A(&L.bqkv, 3 * d); A(&L.bo, d); A(&L.b1, f); A(&L.b2, d);
    A(&L.ln1g, d); A(&L.ln1b, d); A(&L.ln2g, d); A(&L.ln2b, d);
    if (h->nseg == 2) { A(&L.iqkv, 3 * d); A(&L.io, d); A(&L.i1, f); A(&L.i2, d); }
    if (D.arch == PG_ARCH_TRANCEPTION) A(&L.conv_taps, 3 * 4 * 64 * 8);
  }
  if (D.arch == PG_ARCH_MSA) {
    h->col_layers.resize(D.layers);
    for (auto& L : h->col_layers) {
      A(&L.wqkv, static_cast<size_t>(3) * d * d * np); A(&L.wo, static_cast<size_t>(d) * d * np);
      A(&L.bqkv, 3 * d); A(&L.bo, d); A(&L.ln1g, d); A(&L.ln1b, d);
      if (h->nseg == 2) { A(&L.iqkv, 3 * d); A(&L.io, d); }
    }
    A(&h->row_pos, static_cast<size_t>(1024) * d);
  }
  if (D.arch == PG_ARCH_TRANCEPTION) {
    A(&h->qkv2, static_cast<size_t>(h->max_rows) * 3 * d * np);
    A(&h->tok_logp, static_cast<size_t>(h->max_rows));
    A(&h->slopes, D.heads);
  }
  if (h->nseg == 2) A(&h->pack_scratch, static_cast<size_t>(D.layers) * 6);
  A(&h->embed, static_cast<size_t>(D.vocab) * d);
  if (D.arch == PG_ARCH_ESM1B || D.arch == PG_ARCH_MSA) A(&h->pos, static_cast<size_t>(D.max_positions + 2) * d);
  A(&h->lnbg, d); A(&h->lnbb, d); A(&h->lnag, d); A(&h->lnab, d);
  A(&h->hdw, static_cast<size_t>(d) * d); A(&h->hdb, d); A(&h->hlng, d); A(&h->hlnb, d); A(&h->hbias, D.vocab);
  A(&h->x, static_cast<size_t>(h->max_rows) * d);
  A(&h->abuf, static_cast<size_t>(h->max_rows) * d * np);
  A(&h->qkv, static_cast<size_t>(h->max_rows) * 3 * d * np);
  A(&h->fbuf, static_cast<size_t>(h->max_rows) * f * np);
  A(&h->hs_a, static_cast<size_t>(h->head_cap) * d);
  A(&h->hs_b, static_cast<size_t>(h->head_cap) * d);
  A(&h->row_sel, 2 * static_cast<size_t>(h->head_cap));  // second half: zeros (row 0 of every compacted sequence)
  A(&h->xc, static_cast<size_t>(h->head_cap) * d);
  A(&h->cabuf, static_cast<size_t>(h->head_cap) * d * np);
  A(&h->cfbuf, static_cast<size_t>(h->head_cap) * f * np);
  if (h->delta) A(&h->cq, static_cast<size_t>(h->head_cap) * 3 * d * np);
  if (rc) {
    std::string m = h->err;
    pg_destroy(h);
    return set_error(rc, m);
  }
  cudaMemset(h->row_sel, 0, 2 * static_cast<size_t>(h->head_cap) * sizeof(int32_t));
  *out = h;
  return PG_OK;
}

int pg_destroy(pg_handle h) {
  if (!h) return PG_OK;
  cudaSetDevice(h->desc.device);
  for (void* p : h->allocs) cudaFree(p);
  if (h->tied_ws) cudaFree(h->tied_ws);
  if (h->base) cudaFree(h->base);
  delete h;
  return PG_OK;
}

int pg_load_weights(pg_handle h, const pg_tensor* tensors, int32_t n) {
  if (!h || !tensors) return set_error(PG_ERR_ARG, "pg_load_weights: null argument");
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  const pg_model_desc& D = h->desc;
  const int d = D.embed_dim, f = D.ffn_dim, np = h->np;
  std::map<std::string, const pg_tensor*> by_name;
  for (int i = 0; i < n; ++i) by_name[tensors[i].name] = &tensors[i];
  std::string missing;
  auto get = [&](const std::string& name, int64_t r, int64_t c) -> const float* {
    auto it = by_name.find(name);
    if (it == by_name.end()) { missing += name + " "; return nullptr; }
    if (it->second->shape[0] != r || it->second->shape[1] != c) { missing += name + "(shape) "; return nullptr; }
    return static_cast<const float*>(it->second->data);
  };
  auto copy = [&](float* dst, const float* src, size_t cnt, float scale = 1.f) {
    if (!src) return;
    scale_copy_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256>>>(src, dst, static_cast<int>(cnt), scale);
  };
  const bool f8 = h->nseg == 2;
  // inv: where the per-row 1/t_n of this matrix go (PG_PREC_F16F8 only)
  auto pack = [&](__half* dst, float* inv, const float* src, int N, int K, float scale = 1.f) {
    if (!src) return;
    const long long tot = static_cast<long long>(N) * K;
    if (f8) launch_pack_weight_f8(src, N, K, K, 1, scale, h->pack_scratch + (h->pack_used++ % (D.layers * 6)), dst, inv, 0);
    else pack_weight_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256>>>(src, N, K, scale, dst, np);
  };
  auto pack_t = [&](__half* dst, float* inv, const float* src, int N, int K) {
    if (!src) return;
    const long long tot = static_cast<long long>(N) * K;
    if (f8) launch_pack_weight_f8(src, N, K, 1, N, 1.f, h->pack_scratch + (h->pack_used++ % (D.layers * 6)), dst, inv, 0);
    else pack_weight_t_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256>>>(src, N, K, dst, np);
  };
  if (D.arch == PG_ARCH_TRANCEPTION) {
    // HF GPT2-style names with the "transformer." prefix stripped by the host; Conv1D weights are [in, out].
    for (int l = 0; l < D.layers; ++l) {
      Layer& L = h->layers[l];
      const std::string p = "h." + std::to_string(l) + ".";
      pack_t(L.wqkv, L.iqkv, get(p + "attn.c_attn.weight", d, 3 * d), 3 * d, d);
      copy(L.bqkv, get(p + "attn.c_attn.bias", 3 * d, 1), 3 * d);
      pack_t(L.wo, L.io, get(p + "attn.c_proj.weight", d, d), d, d);
      copy(L.bo, get(p + "attn.c_proj.bias", d, 1), d);
      pack_t(L.w1, L.i1, get(p + "mlp.c_fc.weight", d, f), f, d);
      copy(L.b1, get(p + "mlp.c_fc.bias", f, 1), f);
      pack_t(L.w2, L.i2, get(p + "mlp.c_proj.weight", f, d), d, f);
      copy(L.b2, get(p + "mlp.c_proj.bias", d, 1), d);
      copy(L.ln1g, get(p + "ln_1.weight", d, 1), d);
      copy(L.ln1b, get(p + "ln_1.bias", d, 1), d);
      copy(L.ln2g, get(p + "ln_2.weight", d, 1), d);
      copy(L.ln2b, get(p + "ln_2.bias", d, 1), d);
      copy(L.conv_taps, get(p + "attn.conv_taps", 3 * 4 * 64, 8), 3 * 4 * 64 * 8);
    }
    copy(h->embed, get("wte.weight", D.vocab, d), static_cast<size_t>(D.vocab) * d);
    copy(h->lnag, get("ln_f.weight", d, 1), d);
    copy(h->lnab, get("ln_f.bias", d, 1), d);
    copy(h->slopes, get("alibi_slopes", D.heads, 1), D.heads);
    if (!missing.empty()) return fail(h, PG_ERR_ARG, "pg_load_weights: missing or mis-shaped tensors: " + missing);
    PG_CUDA_OK(cudaGetLastError());
    PG_CUDA_OK(cudaDeviceSynchronize());
    h->loaded = true;
    return PG_OK;
  }
  const float qscale = 0.125f;  // head_dim^-1/2 with head_dim == 64 (multihead_attention.py:103,261); exact power of two
  // one attention block: q/k/v/out projections at `pa` ("...self_attn." / "...row_self_attention.layer."), its LayerNorm at `pl`
  auto load_attn = [&](Layer& L, const std::string& pa, const std::string& pl) {
    const size_t dd = static_cast<size_t>(d) * d * np;
    if (f8) {  // q, k, v form ONE GEMM operand: one common e4m3 scale (matrix maximum over the three, q already scaled)
      const float* wq = get(pa + "q_proj.weight", d, d);
      const float* wk = get(pa + "k_proj.weight", d, d);
      const float* wv = get(pa + "v_proj.weight", d, d);
      if (wq && wk && wv) {
        unsigned int* word = h->pack_scratch + (h->pack_used++ % (D.layers * 6));
        const unsigned nb = static_cast<unsigned>((static_cast<long long>(d) * d + 255) / 256);
        cudaMemsetAsync(word, 0, sizeof(unsigned int), 0);
        weight_absmax_kernel<<<nb, 256>>>(wq, d, d, d, 1, qscale, word);
        weight_absmax_kernel<<<nb, 256>>>(wk, d, d, d, 1, 1.f, word);
        weight_absmax_kernel<<<nb, 256>>>(wv, d, d, d, 1, 1.f, word);
        pack_weight_f8_kernel<<<(d + 7) / 8, 256>>>(wq, d, d, d, 1, qscale, word, L.wqkv, L.iqkv);
        pack_weight_f8_kernel<<<(d + 7) / 8, 256>>>(wk, d, d, d, 1, 1.f, word, L.wqkv + dd, L.iqkv + d);
        pack_weight_f8_kernel<<<(d + 7) / 8, 256>>>(wv, d, d, d, 1, 1.f, word, L.wqkv + 2 * dd, L.iqkv + 2 * d);
      }
    } else {
      pack(L.wqkv, nullptr, get(pa + "q_proj.weight", d, d), d, d, qscale);
      pack(L.wqkv + dd, nullptr, get(pa + "k_proj.weight", d, d), d, d);
      pack(L.wqkv + 2 * dd, nullptr, get(pa + "v_proj.weight", d, d), d, d);
    }
    copy(L.bqkv, get(pa + "q_proj.bias", d, 1), d, qscale);
    copy(L.bqkv + d, get(pa + "k_proj.bias", d, 1), d);
    copy(L.bqkv + 2 * d, get(pa + "v_proj.bias", d, 1), d);
    pack(L.wo, L.io, get(pa + "out_proj.weight", d, d), d, d);
    copy(L.bo, get(pa + "out_proj.bias", d, 1), d);
    copy(L.ln1g, get(pl + "weight", d, 1), d);
    copy(L.ln1b, get(pl + "bias", d, 1), d);
  };
  auto load_ffn = [&](Layer& L, const std::string& pf, const std::string& pl) {
    pack(L.w1, L.i1, get(pf + "fc1.weight", f, d), f, d);
    copy(L.b1, get(pf + "fc1.bias", f, 1), f);
    pack(L.w2, L.i2, get(pf + "fc2.weight", d, f), d, f);
    copy(L.b2, get(pf + "fc2.bias", d, 1), d);
    copy(L.ln2g, get(pl + "weight", d, 1), d);
    copy(L.ln2b, get(pl + "bias", d, 1), d);
  };
  for (int l = 0; l < D.layers; ++l) {
    const std::string p = "layers." + std::to_string(l) + ".";
    if (D.arch == PG_ARCH_MSA) {  // AxialTransformerLayer (modules.py:194-196): three NormalizedResidualBlocks {layer, layer_norm}
      load_attn(h->layers[l], p + "row_self_attention.layer.", p + "row_self_attention.layer_norm.");
      load_attn(h->col_layers[l], p + "column_self_attention.layer.", p + "column_self_attention.layer_norm.");
      load_ffn(h->layers[l], p + "feed_forward_layer.layer.", p + "feed_forward_layer.layer_norm.");
    } else {
      load_attn(h->layers[l], p + "self_attn.", p + "self_attn_layer_norm.");
      load_ffn(h->layers[l], p, p + "final_layer_norm.");
    }
  }
  if (D.arch == PG_ARCH_MSA) {
    auto it = by_name.find("msa_position_embedding");  // host-expanded to [1024, d] (the first release stores a 1-wide table)
    if (it != by_name.end()) {
      if (it->second->shape[0] != 1024 || it->second->shape[1] != d) missing += "msa_position_embedding(shape) ";
      else copy(h->row_pos, static_cast<const float*>(it->second->data), static_cast<size_t>(1024) * d);
    } else {
      h->row_pos = nullptr;  // embed_positions_msa off (msa_transformer.py:106-114)
    }
  }
  copy(h->embed, get("embed_tokens.weight", D.vocab, d), static_cast<size_t>(D.vocab) * d);
  if (D.arch == PG_ARCH_ESM1B || D.arch == PG_ARCH_MSA)
    copy(h->pos, get("embed_positions.weight", D.max_positions + 2, d), static_cast<size_t>(D.max_positions + 2) * d);
  if (D.emb_ln_before || D.arch == PG_ARCH_MSA) {
    copy(h->lnbg, get("emb_layer_norm_before.weight", d, 1), d);
    copy(h->lnbb, get("emb_layer_norm_before.bias", d, 1), d);
  }
  copy(h->lnag, get("emb_layer_norm_after.weight", d, 1), d);
  copy(h->lnab, get("emb_layer_norm_after.bias", d, 1), d);
  copy(h->hdw, get("lm_head.dense.weight", d, d), static_cast<size_t>(d) * d);
  copy(h->hdb, get("lm_head.dense.bias", d, 1), d);
  copy(h->hlng, get("lm_head.layer_norm.weight", d, 1), d);
  copy(h->hlnb, get("lm_head.layer_norm.bias", d, 1), d);
  copy(h->hbias, get("lm_head.bias", D.vocab, 1), D.vocab);
  if (D.arch == PG_ARCH_ESM2) {
    auto it = by_name.find("rotary.cos");
    auto is = by_name.find("rotary.sin");
    if (it == by_name.end() || is == by_name.end() || it->second->shape[1] != 32 || is->second->shape[0] != it->second->shape[0]) {
      missing += "rotary.cos/rotary.sin[T,32] ";
    } else {
      h->rot_rows = static_cast<int>(it->second->shape[0]);
      int rc = dev_alloc(h, &h->rot_cos, static_cast<size_t>(h->rot_rows) * 32);
      if (!rc) rc = dev_alloc(h, &h->rot_sin, static_cast<size_t>(h->rot_rows) * 32);
      if (rc) return rc;
      copy(h->rot_cos, static_cast<const float*>(it->second->data), static_cast<size_t>(h->rot_rows) * 32);
      copy(h->rot_sin, static_cast<const float*>(is->second->data), static_cast<size_t>(h->rot_rows) * 32);
    }
  }
  if (!missing.empty()) return fail(h, PG_ERR_ARG, "pg_load_weights: missing or mis-shaped tensors: " + missing);
  PG_CUDA_OK(cudaGetLastError());
  PG_CUDA_OK(cudaDeviceSynchronize());  // caller may free its tensors on return
  h->loaded = true;
  return PG_OK;
}

// Sequences per pass: as many as the workspace holds, then evened out over the passes that needs — 512 copies of a 514-token window
// against the default 131 072-row workspace are 171 + 171 + 170, not 255 + 255 + 2 (a 2-sequence pass pays all 33 layers' launch
// latencies and partial waves for 0.4 % of the rows). Results do not depend on the split (rows are independent; tested).
static long long balanced_per(long long total, long long cap) {
  if (cap < 1) cap = 1;
  if (total <= cap) return total > 0 ? total : 1;
  const long long passes = (total + cap - 1) / cap;
  return (total + passes - 1) / passes;
}

int pg_masked_marginals(pg_handle h, const int32_t* tokens, int32_t n_tokens, const int32_t* positions, const int32_t* win_start,
                        const int32_t* out_row, int32_t P, int32_t T, float* out_logprobs, pg_stream stream) {
  if (!h) return set_error(PG_ERR_ARG, "pg_masked_marginals: null handle");
  if (!h->loaded) return fail(h, PG_ERR_STATE, "pg_masked_marginals: weights not loaded");
  if (h->desc.arch == PG_ARCH_TRANCEPTION) return fail(h, PG_ERR_STATE, "pg_masked_marginals: handle is a Tranception model (use pg_ar_loglik)");
  if (h->desc.arch == PG_ARCH_MSA) return fail(h, PG_ERR_STATE, "pg_masked_marginals: handle is an MSA Transformer (use pg_msa_masked_marginals)");
  if (!tokens || !positions || !out_logprobs) return fail(h, PG_ERR_ARG, "pg_masked_marginals: null buffer");
  if (P < 0 || T <= 0 || T > n_tokens) return fail(h, PG_ERR_ARG, "pg_masked_marginals: bad P/T");
  if (h->desc.arch == PG_ARCH_ESM1B && T > h->desc.max_positions)
    return fail(h, PG_ERR_ARG, "pg_masked_marginals: window longer than the learned position table (modules.py:256-260)");
  if (h->desc.arch == PG_ARCH_ESM2 && T > h->rot_rows) return fail(h, PG_ERR_ARG, "pg_masked_marginals: window longer than rotary tables");
  if (T > h->max_rows) return fail(h, PG_ERR_ARG, "pg_masked_marginals: window longer than workspace (raise max_rows)");
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long per = h->max_rows / T;
  if (per > h->head_cap) per = h->head_cap;
  per = balanced_per(P, per);
  // delta-operand mode needs one shared window (every copy sees all n_tokens tokens) and the masked row as the emitted row
  const bool delta = h->delta && T == n_tokens && out_row == nullptr && P > 0;
  if (delta) {
    int rc = forward_base(h, tokens, n_tokens, T, s);
    if (rc) return fail(h, rc, tls_error());
  }
  for (int p0 = 0; p0 < P; p0 += static_cast<int>(per)) {
    const int Bc = (P - p0) < per ? (P - p0) : static_cast<int>(per);
    row_select_kernel<<<(Bc + 255) / 256, 256, 0, s>>>(positions, win_start, out_row, p0, Bc, h->row_sel);
    int rc = delta ? forward_rows_delta(h, tokens, n_tokens, positions, win_start, p0, Bc, T, s, h->row_sel)
                   : forward_rows(h, tokens, n_tokens, positions, win_start, p0, Bc, T, s, h->row_sel);
    if (rc) return fail(h, rc, tls_error());
    ProfScope ps(CAT_HEAD, s, 4);
    HeadLaunch hl = head_args(h, T);
    hl.x = h->xc; hl.P = Bc; hl.all_rows = 1;  // the pruned last layer left exactly the rows to emit, compacted
    hl.out = out_logprobs + static_cast<long long>(p0) * h->desc.vocab;
    rc = launch_head(hl, s);
    if (rc) return fail(h, rc, tls_error());
  }
  return PG_OK;
}

int pg_msa_masked_marginals(pg_handle h, const int32_t* tokens, int32_t R, int32_t C_full, const int32_t* positions,
                            const int32_t* win_start, int32_t P, int32_t Cw, float* out_logprobs, pg_stream stream) {
  if (!h) return set_error(PG_ERR_ARG, "pg_msa_masked_marginals: null handle");
  if (!h->loaded) return fail(h, PG_ERR_STATE, "pg_msa_masked_marginals: weights not loaded");
  if (h->desc.arch != PG_ARCH_MSA) return fail(h, PG_ERR_STATE, "pg_msa_masked_marginals: handle is not an MSA Transformer");
  if (!tokens || !positions || !out_logprobs) return fail(h, PG_ERR_ARG, "pg_msa_masked_marginals: null buffer");
  if (P < 0 || R <= 0 || Cw <= 0 || Cw > C_full) return fail(h, PG_ERR_ARG, "pg_msa_masked_marginals: bad P/R/Cw");
  if (Cw > h->desc.max_positions) return fail(h, PG_ERR_ARG, "pg_msa_masked_marginals: window longer than the learned position table");
  if (R > 1024 && h->row_pos) return fail(h, PG_ERR_ARG, "pg_msa_masked_marginals: the row-position table covers 1024 alignment rows (msa_transformer.py:163-168)");
  if (Cw < C_full && !win_start) return fail(h, PG_ERR_ARG, "pg_msa_masked_marginals: windows need win_start");
  const long long per_msa = static_cast<long long>(R) * Cw;
  if (per_msa > h->max_rows) return fail(h, PG_ERR_ARG, "pg_msa_masked_marginals: one alignment exceeds the workspace (raise max_rows)");
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long per = h->max_rows / per_msa;
  if (per > h->head_cap) per = h->head_cap;
  per = balanced_per(P, per);
  for (int p0 = 0; p0 < P; p0 += static_cast<int>(per)) {
    const int Bc = (P - p0) < per ? (P - p0) : static_cast<int>(per);
    row_select_kernel<<<(Bc + 255) / 256, 256, 0, s>>>(positions, win_start, nullptr, p0, Bc, h->row_sel);
    int rc = forward_msa(h, tokens, R, C_full, positions, win_start, p0, Bc, Cw, s, h->row_sel);
    if (rc) return fail(h, rc, tls_error());
    ProfScope ps(CAT_HEAD, s, 5);
    HeadLaunch hl = head_args(h, 1);
    hl.x = h->xc; hl.P = Bc; hl.all_rows = 1;
    hl.out = out_logprobs + static_cast<long long>(p0) * h->desc.vocab;
    rc = launch_head(hl, s);
    if (rc) return fail(h, rc, tls_error());
  }
  return PG_OK;
}

int pg_forward_logprobs(pg_handle h, const int32_t* tokens, int32_t n_tokens, int32_t win_start, int32_t T, int32_t mask_pos,
                        float* out_logprobs, pg_stream stream) {
  if (!h) return set_error(PG_ERR_ARG, "pg_forward_logprobs: null handle");
  if (!h->loaded) return fail(h, PG_ERR_STATE, "pg_forward_logprobs: weights not loaded");
  if (h->desc.arch == PG_ARCH_TRANCEPTION) return fail(h, PG_ERR_STATE, "pg_forward_logprobs: handle is a Tranception model (use pg_ar_loglik)");
  if (h->desc.arch == PG_ARCH_MSA) return fail(h, PG_ERR_STATE, "pg_forward_logprobs: handle is an MSA Transformer (use pg_msa_masked_marginals)");
  if (!tokens || !out_logprobs || T <= 0 || win_start < 0 || win_start + T > n_tokens) return fail(h, PG_ERR_ARG, "pg_forward_logprobs: bad arguments");
  if (h->desc.arch == PG_ARCH_ESM1B && T > h->desc.max_positions) return fail(h, PG_ERR_ARG, "pg_forward_logprobs: sequence longer than the learned position table");
  if (h->desc.arch == PG_ARCH_ESM2 && T > h->rot_rows) return fail(h, PG_ERR_ARG, "pg_forward_logprobs: sequence longer than rotary tables");
  if (T > h->max_rows) return fail(h, PG_ERR_ARG, "pg_forward_logprobs: sequence longer than workspace");
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // one row: positions / win_start as 1-element device arrays kept in row_sel scratch
  int32_t hostv[2] = {mask_pos, win_start};
  PG_CUDA_OK(cudaMemcpyAsync(h->row_sel, hostv, sizeof(hostv), cudaMemcpyHostToDevice, s));
  int rc = forward_rows(h, tokens, n_tokens, h->row_sel, h->row_sel + 1, 0, 1, T, s);
  if (rc) return fail(h, rc, tls_error());
  for (int r0 = 0; r0 < T; r0 += h->head_cap) {
    HeadLaunch hl = head_args(h, T);
    hl.x = h->x + static_cast<long long>(r0) * h->desc.embed_dim;
    hl.P = (T - r0) < h->head_cap ? (T - r0) : h->head_cap; hl.all_rows = 1;
    hl.out = out_logprobs + static_cast<long long>(r0) * h->desc.vocab;
    { ProfScope ps(CAT_HEAD, s, 3); rc = launch_head(hl, s); }
    if (rc) return fail(h, rc, tls_error());
  }
  return PG_OK;
}

static int ar_fusion_args(pg_handle h, const pg_ar_fusion* f, ArFusion* base) {
  if (!f) return PG_OK;
  if ((f->log_prior == nullptr) != (f->prior_row == nullptr)) return fail(h, PG_ERR_ARG, "pg_ar_loglik: log_prior and prior_row go together");
  if ((f->log_prior2 == nullptr) != (f->prior_row2 == nullptr)) return fail(h, PG_ERR_ARG, "pg_ar_loglik: log_prior2 and prior_row2 go together");
  if (f->log_prior2 && !f->log_prior) return fail(h, PG_ERR_ARG, "pg_ar_loglik: the second prior needs the first (it is fused inside the MSA overlap)");
  if (f->first_col < 0 || f->first_col > h->desc.vocab) return fail(h, PG_ERR_ARG, "pg_ar_loglik: first_col outside the vocabulary");
  base->log_prior = f->log_prior; base->prior_row = f->prior_row; base->alpha = f->alpha;
  base->log_prior2 = f->log_prior2; base->prior_row2 = f->prior_row2; base->beta = f->beta;
  base->first_col = f->first_col; base->out_logprobs = f->out_logprobs;
  return PG_OK;
}

int pg_ar_loglik_fused(pg_handle h, const int32_t* ids, const int32_t* lens, int32_t B, int32_t T, const pg_ar_fusion* f,
                       float* out_sum_logp, pg_stream stream) {
  if (!h) return set_error(PG_ERR_ARG, "pg_ar_loglik: null handle");
  if (!h->loaded) return fail(h, PG_ERR_STATE, "pg_ar_loglik: weights not loaded");
  if (h->desc.arch != PG_ARCH_TRANCEPTION) return fail(h, PG_ERR_STATE, "pg_ar_loglik: handle is not a Tranception model");
  if (B < 0 || T <= 0) return fail(h, PG_ERR_ARG, "pg_ar_loglik: bad B/T");
  if (B == 0) return PG_OK;
  if (!ids || !lens || !out_sum_logp) return fail(h, PG_ERR_ARG, "pg_ar_loglik: null buffer");
  ArFusion base;
  { int frc = ar_fusion_args(h, f, &base); if (frc) return frc; }
  if (T > h->desc.max_positions) return fail(h, PG_ERR_ARG, "pg_ar_loglik: sequence longer than n_ctx");
  if (T > h->max_rows) return fail(h, PG_ERR_ARG, "pg_ar_loglik: sequence longer than workspace");
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long per = balanced_per(B, h->max_rows / T);
  for (int b0 = 0; b0 < B; b0 += static_cast<int>(per)) {
    const int Bc = (B - b0) < per ? (B - b0) : static_cast<int>(per);
    const long long off = static_cast<long long>(b0) * T;
    const int32_t* idc = ids + off;
    int rc = forward_tranception(h, idc, Bc, T, s);
    if (rc) return fail(h, rc, tls_error());
    ProfScope ps(CAT_HEAD, s, 2);
    ArFusion fc = base;
    if (fc.prior_row) fc.prior_row += off;
    if (fc.prior_row2) fc.prior_row2 += off;
    if (fc.out_logprobs) fc.out_logprobs += off * h->desc.vocab;
    rc = launch_ar_head(h->x, h->desc.embed_dim, Bc, T, h->desc.vocab, idc, lens + b0, h->lnag, h->lnab, h->embed, fc, h->tok_logp,
                        out_sum_logp + b0, s);
    if (rc) return fail(h, rc, tls_error());
  }
  return PG_OK;
}

int pg_ar_prefix_begin(pg_handle h, const int32_t* ids, int32_t T, const pg_ar_fusion* f, float* out_tok_logp, pg_stream stream) {
  if (!h) return set_error(PG_ERR_ARG, "pg_ar_prefix_begin: null handle");
  if (!h->loaded) return fail(h, PG_ERR_STATE, "pg_ar_prefix_begin: weights not loaded");
  if (h->desc.arch != PG_ARCH_TRANCEPTION) return fail(h, PG_ERR_STATE, "pg_ar_prefix_begin: handle is not a Tranception model");
  if (!ids || !out_tok_logp || T < 2) return fail(h, PG_ERR_ARG, "pg_ar_prefix_begin: bad arguments");
  if (T > h->desc.max_positions || T > h->max_rows) return fail(h, PG_ERR_ARG, "pg_ar_prefix_begin: sequence longer than n_ctx / workspace");
  ArFusion fu;
  int rc = ar_fusion_args(h, f, &fu);
  if (rc) return rc;
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t ldq = static_cast<size_t>(3) * h->desc.embed_dim * h->np;
  if (h->prefix_cap < h->desc.max_positions) {  // first use: [layers][n_ctx][3d*np] fp16, twice
    const size_t n = static_cast<size_t>(h->desc.layers) * h->desc.max_positions * ldq;
    rc = dev_alloc(h, &h->raw_cache, n);
    if (!rc) rc = dev_alloc(h, &h->kv_cache, n);
    if (rc) return rc;
    h->prefix_cap = h->desc.max_positions;
  }
  h->prefix_T = 0;
  rc = forward_tranception(h, ids, 1, T, s, 1, 0);
  if (rc) return fail(h, rc, tls_error());
  // per-token log p(ids[t+1] | ids[<=t]) of the wild type, t = 0 .. T-2 (and the fused rows in f->out_logprobs when asked for)
  int32_t* len_dev = h->row_sel;
  PG_CUDA_OK(cudaMemcpyAsync(len_dev, &T, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  { ProfScope ps(CAT_HEAD, s, 2);
    rc = launch_ar_head(h->x, h->desc.embed_dim, 1, T, h->desc.vocab, ids, len_dev, h->lnag, h->lnab, h->embed, fu, out_tok_logp, h->tok_logp, s); }
  if (rc) return fail(h, rc, tls_error());
  h->prefix_T = T;
  return PG_OK;
}

int pg_ar_loglik_prefix(pg_handle h, const int32_t* ids, const int32_t* lens, int32_t B, int32_t T, int32_t start, const pg_ar_fusion* f,
                        float* out_sum_logp, pg_stream stream) {
  if (!h) return set_error(PG_ERR_ARG, "pg_ar_loglik_prefix: null handle");
  if (!h->loaded) return fail(h, PG_ERR_STATE, "pg_ar_loglik_prefix: weights not loaded");
  if (h->desc.arch != PG_ARCH_TRANCEPTION) return fail(h, PG_ERR_STATE, "pg_ar_loglik_prefix: handle is not a Tranception model");
  if (h->prefix_T <= 0) return fail(h, PG_ERR_STATE, "pg_ar_loglik_prefix: no wild-type prefix recorded (call pg_ar_prefix_begin first)");
  if (start <= 0 || start % 128 || start >= h->prefix_T) return fail(h, PG_ERR_ARG, "pg_ar_loglik_prefix: start must be a positive multiple of 128 below the recorded length");
  if (B < 0 || T <= 0) return fail(h, PG_ERR_ARG, "pg_ar_loglik_prefix: bad B/T");
  if (B == 0) return PG_OK;
  if (!ids || !lens || !out_sum_logp) return fail(h, PG_ERR_ARG, "pg_ar_loglik_prefix: null buffer");
  if (start + T > h->desc.max_positions) return fail(h, PG_ERR_ARG, "pg_ar_loglik_prefix: sequence longer than n_ctx");
  if (T > h->max_rows) return fail(h, PG_ERR_ARG, "pg_ar_loglik_prefix: suffix longer than workspace");
  ArFusion base;
  int rc = ar_fusion_args(h, f, &base);
  if (rc) return rc;
  PG_CUDA_OK(cudaSetDevice(h->desc.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long per = balanced_per(B, h->max_rows / T);
  for (int b0 = 0; b0 < B; b0 += static_cast<int>(per)) {
    const int Bc = (B - b0) < per ? (B - b0) : static_cast<int>(per);
    const long long off = static_cast<long long>(b0) * T;
    rc = forward_tranception(h, ids + off, Bc, T, s, 2, start);
    if (rc) return fail(h, rc, tls_error());
    ProfScope ps(CAT_HEAD, s, 2);
    ArFusion fc = base;
    if (fc.prior_row) fc.prior_row += off;
    if (fc.prior_row2) fc.prior_row2 += off;
    if (fc.out_logprobs) fc.out_logprobs += off * h->desc.vocab;
    rc = launch_ar_head(h->x, h->desc.embed_dim, Bc, T, h->desc.vocab, ids + off, lens + b0, h->lnag, h->lnab, h->embed, fc, h->tok_logp,
                        out_sum_logp + b0, s);
    if (rc) return fail(h, rc, tls_error());
  }
  return PG_OK;
}

int pg_ar_loglik(pg_handle h, const int32_t* ids, const int32_t* lens, int32_t B, int32_t T, const float* log_prior,
                 const int32_t* prior_row, float alpha, float* out_sum_logp, pg_stream stream) {
  pg_ar_fusion f = {};
  f.log_prior = log_prior; f.prior_row = prior_row; f.alpha = alpha;
  return pg_ar_loglik_fused(h, ids, lens, B, T, &f, out_sum_logp, stream);
}

int pg_score_mutants(const float* table, int32_t n_rows, int32_t vocab, const int32_t* site_row, const int32_t* site_wt,
                     const int32_t* site_mt, const int32_t* row_offsets, int32_t M, float* out_scores, pg_stream stream) {
  if (M < 0 || n_rows < 0 || vocab <= 0) return set_error(PG_ERR_ARG, "pg_score_mutants: bad sizes");
  if (M == 0) return PG_OK;
  if (!table || !site_row || !site_wt || !site_mt || !row_offsets || !out_scores) return set_error(PG_ERR_ARG, "pg_score_mutants: null buffer");
  ProfScope ps(CAT_SCORE, static_cast<cudaStream_t>(stream));
  return launch_score(table, n_rows, vocab, site_row, site_wt, site_mt, row_offsets, M, out_scores, static_cast<cudaStream_t>(stream));
}

int pg_set_tuning(const char* key, int32_t value) {
  if (!key) return set_error(PG_ERR_ARG, "pg_set_tuning: null key");
  if (std::string(key) == "gemm_kchunk") { set_gemm_kchunk(value); return PG_OK; }
  if (std::string(key) == "gemm_prefetch") { set_gemm_prefetch(value); return PG_OK; }
  if (std::string(key) == "gemm_cta2") { set_gemm_cta2(value); return PG_OK; }
  return set_error(PG_ERR_ARG, std::string("pg_set_tuning: unknown key ") + key);
}

int pg_gemm(const pg_gemm_args* a, pg_stream stream) {
  if (!a) return set_error(PG_ERR_ARG, "pg_gemm: null args");
  GemmLaunch g{};
  g.a = a->a; g.lda = a->lda; g.w = a->w; g.ldw = a->ldw; g.bias = a->bias;
  g.M = a->M; g.N = a->N; g.K = a->K; g.nseg = a->nseg; g.epi = a->epi;
  g.out = static_cast<__half*>(a->out_h); g.ldo = a->ldo; g.out_lo_off = a->out_lo_off;
  g.resid = a->resid; g.ldr = a->ldr;
  g.rot_cos = a->rot_cos; g.rot_sin = a->rot_sin; g.rot_T = a->rot_T; g.rot_dim = a->rot_dim;
  g.a_scale = a->a_scale; g.w_inv = a->w_inv; g.out_scale = a->out_scale;
  g.out_fmt = a->out_fmt ? a->out_fmt : (a->out_lo_off > 0 ? 1 : 0);
  g.grp_rows_a = a->grp_rows_a; g.grp_rows_b = a->grp_rows_b;
  g.base_pre = a->base_pre; g.base_post = static_cast<const __half*>(a->base_post); g.base_T = a->base_T; g.mask_pos = a->mask_pos;
  ProfScope ps(CAT_OTHER, static_cast<cudaStream_t>(stream));
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int pg_layernorm_f16(const float* x, int64_t ldx, const float* gamma, const float* beta, int32_t rows, int32_t d, void* out,
                     int64_t ldo, int64_t lo_off, int32_t fmt, float scale, pg_stream stream) {
  if (!x || !gamma || !beta || !out) return set_error(PG_ERR_ARG, "pg_layernorm_f16: null buffer");
  return launch_layernorm_f16(x, ldx, gamma, beta, rows, d, static_cast<__half*>(out), ldo, lo_off, static_cast<cudaStream_t>(stream),
                              fmt ? fmt : -1, scale);
}

int pg_pack_weight(const float* w, int32_t N, int32_t K, int32_t fmt, void* out, float* w_inv, pg_stream stream) {
  if (!w || !out || N <= 0 || K <= 0) return set_error(PG_ERR_ARG, "pg_pack_weight: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long tot = static_cast<long long>(N) * K;
  if (fmt == 2) {
    if (!w_inv) return set_error(PG_ERR_ARG, "pg_pack_weight: fmt 2 needs w_inv[N]");
    // the matrix maximum is accumulated in w_inv[0]'s storage (as bits) before the pack kernel overwrites it with 1 / t... a separate
    // word is cleaner: the last row's slot is written last by its own warp only, so use a small static scratch per device instead
    static unsigned int* scratch[64] = {};
    int dev = 0;
    PG_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 64) return set_error(PG_ERR_UNSUPPORTED, "pg_pack_weight: device ordinal >= 64");
    if (!scratch[dev]) PG_CUDA_OK(cudaMalloc(&scratch[dev], 256));
    launch_pack_weight_f8(w, N, K, K, 1, 1.f, scratch[dev], static_cast<__half*>(out), w_inv, s);
  } else if (fmt == 0 || fmt == 1) {
    pack_weight_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(w, N, K, 1.f, static_cast<__half*>(out), fmt + 1);
  } else {
    return set_error(PG_ERR_ARG, "pg_pack_weight: fmt must be 0, 1 or 2");
  }
  PG_CUDA_OK(cudaGetLastError());
  return PG_OK;
}

int pg_attention(const pg_attn_args* a, pg_stream stream) {
  if (!a || !a->qkv || !a->out) return set_error(PG_ERR_ARG, "pg_attention: null args");
  AttnLaunch l{};
  l.qkv = static_cast<const __half*>(a->qkv); l.ld = a->ld; l.lo_off = a->lo_off;
  l.out = static_cast<__half*>(a->out); l.ldo = a->ldo; l.out_lo_off = a->out_lo_off;
  l.B = a->B; l.T = a->T; l.heads = a->heads; l.nseg = a->nseg; l.causal = a->causal; l.alibi_slopes = a->alibi_slopes;
  l.out_fmt = a->out_fmt ? a->out_fmt : -1; l.out_scale = a->out_scale;
  l.base_o = static_cast<const __half*>(a->base_o); l.mask_pos = a->mask_pos; l.cout = static_cast<__half*>(a->cout); l.ldc = a->ldc; l.c_lo_off = a->c_lo_off;
  if ((l.base_o || l.mask_pos) && a->impl != 0) return set_error(PG_ERR_UNSUPPORTED, "pg_attention: the delta-operand form needs impl 0");
  ProfScope ps(CAT_OTHER, static_cast<cudaStream_t>(stream));
  if (a->impl == 0) return launch_attention_tc(l, static_cast<cudaStream_t>(stream));   // the model's kernel (tcgen05)
  if (a->impl != 1) return set_error(PG_ERR_ARG, "pg_attention: impl must be 0 (the model's tcgen05 kernel) or 1 (mma.sync cross-check)");
  if (l.out_fmt == 2) return set_error(PG_ERR_UNSUPPORTED, "pg_attention: the mma.sync cross-check kernel writes fp16 planes only");
  return launch_attention(l, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
