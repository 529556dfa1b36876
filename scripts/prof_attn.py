import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proteingym_b200 import _lib
lib = _lib.load()
B, T, H = 64, 514, 20; d = H * 64
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
np_ = 2 if nseg == 3 else 1
x32 = torch.randn(B * T, 3 * d, device="cuda") * 0.5
hi = x32.half()
qkv = torch.cat([hi, (x32 - hi.float()).half()], 1).contiguous() if np_ == 2 else hi   # lo plane = the fp16 remainder, as in the model
out = torch.empty(B * T, d * np_, device="cuda", dtype=torch.float16)
a = _lib.PgAttnArgs(); a.qkv = qkv.data_ptr(); a.ld = 3 * d * np_; a.lo_off = 3 * d if nseg == 3 else 0
a.out = out.data_ptr(); a.ldo = d * np_; a.out_lo_off = d if nseg == 3 else 0
a.B, a.T, a.heads, a.nseg, a.causal, a.impl = B, T, H, nseg, 0, (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for _ in range(3): lib.pg_attention(C.byref(a), None)
torch.cuda.synchronize()
