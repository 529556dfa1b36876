"""Launch every kernel of the ESM / Tranception / MSA paths once at production shape (ESM-1v 650M, 128 sequences x 514 tokens)
so one `ncu --set full` run captures them all:  ncu --set full --clock-control none --import-source on -o out python scripts/prof_kernels.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from proteingym_b200 import _lib
lib = _lib.load()
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
np_ = 2 if nseg == 3 else 1
B, T, H, d, f = 128, 514, 20, 1280, 5120
rows = B * T
x = torch.randn(rows, d, device="cuda"); g = torch.ones(d, device="cuda"); b = torch.zeros(d, device="cuda")
ab = torch.empty(rows, d * np_, device="cuda", dtype=torch.float16)
lib.pg_layernorm_f16(x.data_ptr(), d, g.data_ptr(), b.data_ptr(), rows, d, ab.data_ptr(), d * np_, d if np_ == 2 else 0, None)
def gemm(M, N, K, epi, a, w, bias, out=None, res=None):
    A = _lib.PgGemmArgs(); A.a = a.data_ptr(); A.lda = K * np_; A.w = w.data_ptr(); A.ldw = K * np_; A.bias = bias.data_ptr()
    A.M, A.N, A.K, A.nseg, A.epi = M, N, K, nseg, epi
    if res is not None: A.resid = res.data_ptr(); A.ldr = N
    else: A.out_h = out.data_ptr(); A.ldo = N * np_; A.out_lo_off = N if np_ == 2 else 0
    assert lib.pg_gemm(C.byref(A), None) == 0
wq = (torch.randn(3 * d, d * np_, device="cuda") * 0.03).half(); bq = torch.zeros(3 * d, device="cuda")
qkv = torch.empty(rows, 3 * d * np_, device="cuda", dtype=torch.float16)
gemm(rows, 3 * d, d, 0, ab, wq, bq, out=qkv)                                    # QKV
at = _lib.PgAttnArgs(); at.qkv = qkv.data_ptr(); at.ld = 3 * d * np_; at.lo_off = 3 * d if np_ == 2 else 0
at.out = ab.data_ptr(); at.ldo = d * np_; at.out_lo_off = d if np_ == 2 else 0; at.B, at.T, at.heads, at.nseg, at.causal, at.impl = B, T, H, nseg, 0, 2
assert lib.pg_attention(C.byref(at), None) == 0                                  # tcgen05 attention
wo = (torch.randn(d, d * np_, device="cuda") * 0.03).half(); bo = torch.zeros(d, device="cuda")
gemm(rows, d, d, 2, ab, wo, bo, res=x)                                           # out_proj + residual
w1 = (torch.randn(f, d * np_, device="cuda") * 0.03).half(); b1 = torch.zeros(f, device="cuda"); fb = torch.empty(rows, f * np_, device="cuda", dtype=torch.float16)
gemm(rows, f, d, 1, ab, w1, b1, out=fb)                                          # fc1 + GELU
w2 = (torch.randn(d, f * np_, device="cuda") * 0.02).half()
gemm(rows, d, f, 2, fb, w2, bo, res=x)                                           # fc2 + residual
at.causal, at.impl = 1, 1; sl = torch.full((H,), 0.01, device="cuda"); at.alibi_slopes = sl.data_ptr()
assert lib.pg_attention(C.byref(at), None) == 0                                  # mma.sync causal + ALiBi (Tranception)
N, L = 20000, 500
mat = torch.randint(0, 21, (N, L), device="cuda", dtype=torch.uint8); need = torch.full((N,), 400, dtype=torch.int32, device="cuda"); out = torch.empty(N, dtype=torch.int32, device="cuda")
lib.pg_msa_cluster_neighbors(mat.data_ptr(), L, N, L, need.data_ptr(), out.data_ptr(), None)
tt = mat.T.contiguous(); w = torch.ones(N, dtype=torch.float64, device="cuda"); po = torch.empty(L, 25, dtype=torch.float64, device="cuda")
lib.pg_msa_prior(tt.data_ptr(), w.data_ptr(), N, L, 25, 1e-5, po.data_ptr(), None)
torch.cuda.synchronize()
