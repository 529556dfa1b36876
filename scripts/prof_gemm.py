import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proteingym_b200 import _lib
lib = _lib.load()
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M, N, K = 65536, 5120, 1280  # fc1 at production shape (128 sequences x 512 rows)
np_ = 2 if nseg == 3 else 1
a = (torch.randn(M, K * np_, device="cuda") * 0.5).half(); w = (torch.randn(N, K * np_, device="cuda") * 0.03).half(); b = torch.zeros(N, device="cuda")
out = torch.empty(M, N * np_, device="cuda", dtype=torch.float16)
g = _lib.PgGemmArgs(); g.a = a.data_ptr(); g.lda = K * np_; g.w = w.data_ptr(); g.ldw = K * np_; g.bias = b.data_ptr()
g.M, g.N, g.K, g.nseg, g.epi = M, N, K, nseg, 1; g.out_h = out.data_ptr(); g.ldo = N * np_; g.out_lo_off = N if nseg == 3 else 0
for _ in range(3): lib.pg_gemm(C.byref(g), None)
torch.cuda.synchronize()
