"""Numerics study (CPU, fp32/fp64 oracle) of a delta-operand scheme for the linear layers of masked-marginal scoring.

The P masked copies of one assay differ from the unmasked sequence x0 by a perturbation: a = a0 + D. A linear layer is then
W a = W a0 + W D with the shared W a0 computed once at full precision and W D per copy. If |D| << |a|, W D tolerates coarser
operands: a single fp16 pass on D has absolute error ~2^-11 |D| |W|, which matches the 2-unit fp16+e4m3 scheme (~2^-15 |a| |W|)
wherever |D| <~ 2^-4 |a|. The row holding the mask (|D| ~ |a|) is computed exactly (it would be one extra small GEMM of P rows).
This script measures |D|/|a| per layer and the end-to-end score error of the scheme at true ESM-1v 650M size with the repo's
synthetic checkpoint (the one the goldens use). Usage: python scripts/precision_delta.py [L] [npos]"""
import sys
sys.path.insert(0, "/root/repo")
import torch
from oracle import esm_oracle as O
from proteingym_b200 import synth
torch.set_num_threads(16)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 96
npos = int(sys.argv[2]) if len(sys.argv) > 2 else 6
arch = synth.ESM1V_650M
st = synth.make_esm_state(arch, seed=0)
seq = synth.random_protein(L, 7)
T = L + 2
pos = list(range(1, L + 1, max(1, L // npos)))[:npos]
f16 = lambda t: t.to(torch.float16).to(t.dtype)
toks = O.tokenize(seq)[None].repeat(len(pos), 1)
for r, i in enumerate(pos):
    toks[r, i] = 32
wtid = torch.tensor([O.TOK[seq[i - 1]] for i in pos])
aa = torch.tensor([O.TOK[a] for a in synth.AA20])


def scores(lp):
    rows = torch.stack([lp[r, i] for r, i in enumerate(pos)]).double()
    return rows[:, aa] - rows[torch.arange(len(pos)), wtid][:, None]


def run(dt, mm=None, rnd=None):
    s = O.load_state(st, "esm1v", dt)
    with torch.no_grad():
        return scores(torch.log_softmax(O.esm_forward(s, toks, "esm1v", arch.layers, arch.heads, True, dt, rnd, mm), -1))


ref = run(torch.float64)
print(f"L={L} positions={pos}", flush=True)
e = (run(torch.float32, rnd=lambda t: f16(t) if t.ndim <= 3 else t) - ref).abs()
print(f"plain single fp16 pass (linear layers only)   max={e.max().item():.2e} mean={e.mean().item():.2e}", flush=True)

# base pass: the unmasked sequence with the masked copies' token-dropout scale 0.88 / (1 - 1/T)  (esm1.py:128-131)
s32 = O.load_state(st, "esm1v", torch.float32)
sb = dict(s32)
sb["lm_head.weight"] = s32["lm_head.weight"] * (0.88 / (1.0 - 1.0 / T))
A0, Y0 = [], []


def mm_base(a, w, site):
    A0.append(a.clone())
    y = (a.double() @ w.double().T).float()
    Y0.append(y)
    return y


with torch.no_grad():
    O.esm_forward(sb, O.tokenize(seq)[None], "esm1v", arch.layers, arch.heads, False, torch.float32, None, mm_base)

ratios = {}
for mode in ("exact-row", "all-rows"):
    k = [0]

    def mm_delta(a, w, site, mode=mode):
        a0, y0 = A0[k[0]], Y0[k[0]]
        k[0] += 1
        D = a - a0
        y = y0 + f16(D) @ f16(w).T
        if mode == "exact-row":
            for r, i in enumerate(pos):
                y[r, i] = a[r, i] @ w.T
        nm = torch.ones(a.shape[:2], dtype=torch.bool)
        for r, i in enumerate(pos):
            nm[r, i] = False
        ratios.setdefault(site, []).append((D[nm].norm() / a[nm].norm()).item())
        return y

    e = (run(torch.float32, mm=mm_delta) - ref).abs()
    print(f"delta scheme, fp16 single pass on D ({mode:9s}) max={e.max().item():.2e} mean={e.mean().item():.2e}", flush=True)
    if mode == "exact-row":
        for site in ("qkv", "out", "fc1", "fc2"):
            v = ratios[site][::3] if site == "qkv" else ratios[site]
            print(f"  |D|/|a| over unmasked rows, {site:3s}: layer 0 {v[0]:.3f}  8 {v[8]:.3f}  16 {v[16]:.3f}  24 {v[24]:.3f}  32 {v[32]:.3f}  max {max(v):.3f}")
        ratios.clear()
