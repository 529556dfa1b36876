"""Probe the fp32 accumulation behaviour of tcgen05.mma kind::f16 (round-to-nearest vs truncation): all-positive products make
a rounding-mode bias visible as a systematic relative error that grows with K."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proteingym_b200 import _lib
lib = _lib.load()
for K in (256, 1280, 5120, 10240, 16384):
    M, N = 256, 256
    g = torch.Generator(device="cuda").manual_seed(K)
    A = (torch.rand(M, K, device="cuda", generator=g) + 0.5).half(); W = (torch.rand(N, K, device="cuda", generator=g) + 0.5).half()
    ref = A.double() @ W.double().T
    res = torch.zeros(M, N, device="cuda"); bias = torch.zeros(N, device="cuda")
    a = _lib.PgGemmArgs(); a.a = A.data_ptr(); a.lda = K; a.w = W.data_ptr(); a.ldw = K; a.bias = bias.data_ptr()
    a.M, a.N, a.K, a.nseg, a.epi = M, N, K, 1, 2; a.resid = res.data_ptr(); a.ldr = N
    lib.pg_gemm(C.byref(a), None); torch.cuda.synchronize()
    rel = ((res.double() - ref) / ref)
    cub = (A.float() @ W.float().T).double()  # cuBLAS fp32 (no tensor cores unless TF32 allowed; default off)
    relc = ((cub - ref) / ref)
    h = (A @ W.T).double()  # cuBLAS fp16 tensor-core GEMM with fp32 accumulate, fp16 output
    print(f"K={K:6d} tcgen05: mean rel err {rel.mean().item():+.3e} (abs mean {rel.abs().mean().item():.3e}, max {rel.abs().max().item():.3e}) | "
          f"cuBLAS fp32: mean {relc.mean().item():+.3e} abs {relc.abs().mean().item():.3e} | K/16*2^-24 = {K/16*2**-24:.3e}")
