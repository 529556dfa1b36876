"""One GEMM shape under ncu: python scripts/prof_gemm.py <nseg 1|2|3> <fc1|fc2|qkv|out>  (3 launches)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proteingym_b200 import _lib
lib = _lib.load()
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "fc1"
M = 65536
N, K, epi = {"qkv": (3840, 1280, 0), "out": (1280, 1280, 2), "fc1": (5120, 1280, 1), "fc2": (1280, 5120, 2)}[which]
np_ = 1 if nseg == 1 else 2
fmt = {1: 0, 3: 1, 2: 2}[nseg]
x = torch.randn(M, K, device="cuda"); g = torch.ones(K, device="cuda"); b = torch.zeros(K, device="cuda")
a = torch.empty(M, K * np_, device="cuda", dtype=torch.float16)
lib.pg_layernorm_f16(x.data_ptr(), K, g.data_ptr(), b.data_ptr(), M, K, a.data_ptr(), K * np_, K if np_ == 2 else 0, fmt, 4.0, None)
w32 = torch.randn(N, K, device="cuda") / K ** 0.5
w = torch.empty(N, K * np_, device="cuda", dtype=torch.float16); winv = torch.ones(N, device="cuda")
lib.pg_pack_weight(w32.data_ptr(), N, K, fmt, w.data_ptr(), winv.data_ptr(), None)
bias = torch.randn(N, device="cuda")
args = _lib.PgGemmArgs()
args.a, args.lda, args.w, args.ldw, args.bias = a.data_ptr(), K * np_, w.data_ptr(), K * np_, bias.data_ptr()
args.M, args.N, args.K, args.nseg, args.epi = M, N, K, nseg, epi
args.a_scale, args.w_inv = 4.0, winv.data_ptr()
if epi == 2:
    res = torch.zeros(M, N, device="cuda"); args.resid, args.ldr = res.data_ptr(), N
else:
    out = torch.empty(M, N * np_, device="cuda", dtype=torch.float16)
    args.out_h, args.ldo, args.out_lo_off = out.data_ptr(), N * np_, (N if np_ == 2 else 0)
    args.out_fmt, args.out_scale = ((2 if nseg == 2 else 0) if which == "fc1" else 0), 2.0
for _ in range(3):
    _lib.check(lib.pg_gemm(C.byref(args), None))
torch.cuda.synchronize()
