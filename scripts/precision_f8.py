"""CPU emulation of the `f16f8` operand scheme (fp16 hi*hi + e4m3 cross terms) at true ESM-1v 650M size, next to f16x3.

    a @ w.T  ~=  hi(a) hi(w)^T  +  q8(lo(a) 2^11 sA) q8(hi(w) t_n)^T / (2^11 sA t_n)  +  q8(hi(a) sA) q8(lo(w) 2^11 t_n)^T / (2^11 sA t_n)

hi = rn_fp16(x), lo = x - hi (exact in fp32), q8 = round-to-nearest e4m3 with saturation at +-448, sA a fixed power of two per
GEMM site (what the producing kernel applies), t_n a power of two per weight row chosen at load time so that the row's largest
|hi| lands in [112, 224]. Attention operands keep the fp16 hi+lo split (three products), as in the CUDA path.
Usage: python scripts/precision_f8.py [L] [npos]   (per-mutant score error vs an fp64 run of the oracle)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import esm_oracle as O
from proteingym_b200 import synth

torch.set_num_threads(8)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 96
npos = int(sys.argv[2]) if len(sys.argv) > 2 else 6
arch = synth.ESM1V_650M
st = synth.make_esm_state(arch, seed=0)
seq = synth.random_protein(L, 7)
pos = list(range(1, L + 1, max(1, L // npos)))[:npos]


def f16(t):
    return t.to(torch.float16).to(t.dtype)


def split2(t):
    hi = f16(t)
    return hi + f16(t - hi)


def q8(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(t.dtype)


def q8_e5m2(t):
    return t.clamp(-57344.0, 57344.0).to(torch.float8_e5m2).to(t.dtype)


def row_scale(w_hi):
    amax = w_hi.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    return torch.exp2(torch.floor(torch.log2(224.0 / amax)))


def make_mm(site_scale, quant=q8, lo_shift=11):
    cache = {}

    def mm(a, w, site):
        key = (w.data_ptr(), site)
        if key not in cache:
            wh = f16(w)
            wl = w - wh
            t = row_scale(wh)
            cache[key] = (wh, quant(wh * t) / t, quant(wl * t * 2.0 ** lo_shift) / (t * 2.0 ** lo_shift))
        wh, wh8, wl8 = cache[key]
        sA = site_scale[site]
        ah = f16(a)
        al = a - ah
        ah8 = quant(ah * sA) / sA
        al8 = quant(al * sA * 2.0 ** lo_shift) / (sA * 2.0 ** lo_shift)
        return ah @ wh.T + al8 @ wh8.T + ah8 @ wl8.T

    return mm


toks = O.tokenize(seq)[None].repeat(len(pos), 1)
for r, i in enumerate(pos):
    toks[r, i] = 32
wtid = torch.tensor([O.TOK[seq[i - 1]] for i in pos])
aa = torch.tensor([O.TOK[a] for a in synth.AA20])


def run(dt, rnd=None, mm=None):
    s = O.load_state(st, "esm1v", dt)
    with torch.no_grad():
        lp = torch.log_softmax(O.esm_forward(s, toks, "esm1v", arch.layers, arch.heads, True, dt, rnd, mm), -1)
    rows = torch.stack([lp[r, i] for r, i in enumerate(pos)]).double()
    return rows[:, aa] - rows[torch.arange(len(pos)), wtid][:, None]


t0 = time.time()
ref = run(torch.float64)
print(f"fp64 reference: {time.time() - t0:.1f} s, L={L}, positions={pos}", flush=True)


def attn_only(t):  # rnd hook now only sees the attention operands (q, k, v, P): fp16 hi+lo
    return split2(t)


cases = [("f16x3 (all operands hi+lo)", dict(rnd=split2)),
         ("f16 single pass", dict(rnd=f16))]
for sA in (1.0, 4.0, 16.0):
    cases.append((f"f16f8 e4m3, sA={sA:g} all sites", dict(rnd=attn_only, mm=make_mm(dict(qkv=sA, out=sA, fc1=sA, fc2=sA)))))
cases.append(("f16f8 e4m3, sA=4, lo_shift 10", dict(rnd=attn_only, mm=make_mm(dict(qkv=4.0, out=4.0, fc1=4.0, fc2=4.0), lo_shift=10))))
cases.append(("f16f8 e5m2, sA=4", dict(rnd=attn_only, mm=make_mm(dict(qkv=4.0, out=4.0, fc1=4.0, fc2=4.0), quant=q8_e5m2))))
for name, kw in cases:
    t0 = time.time()
    e = (run(torch.float32, **kw) - ref).abs()
    print(f"{name:36s} max={e.max().item():.2e} mean={e.mean().item():.2e}  ({time.time() - t0:.0f} s)", flush=True)
