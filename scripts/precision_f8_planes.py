"""Which of the four e4m3 planes of the f16f8 scheme costs what: per-plane ablation at true ESM-1v 650M size on CPU (see scripts/precision_f8.py
for the scheme). Result (r02, L=96, 6 positions): every plane alone 3.4-7.3e-5 mean / 1.3-2.4e-4 max, all four 8.9e-5 / 2.7e-4: no single
extra digit halves the error."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import esm_oracle as O
from proteingym_b200 import synth
torch.set_num_threads(8)
L, npos = 96, 6
arch = synth.ESM1V_650M
st = synth.make_esm_state(arch, seed=0)
seq = synth.random_protein(L, 7)
pos = list(range(1, L + 1, max(1, L // npos)))[:npos]
f16 = lambda t: t.to(torch.float16).to(t.dtype)
def split2(t):
    hi = f16(t); return hi + f16(t - hi)
def q8(t): return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(t.dtype)
ident = lambda t: t
def row_scale(w_hi):
    amax = w_hi.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    return torch.exp2(torch.floor(torch.log2(224.0 / amax)))
def q2(t):  # two e4m3 digits
    a = q8(t); return a + q8((t - a) * 16.0) / 16.0
def make_mm(qa_lo, qa_hi, qw_hi, qw_lo, sA=4.0):
    cache = {}
    def mm(a, w, site):
        key = (w.data_ptr(), site)
        if key not in cache:
            wh = f16(w); wl = w - wh; t = row_scale(wh)
            cache[key] = (wh, qw_hi(wh * t) / t, qw_lo(wl * t * 2048.0) / (t * 2048.0))
        wh, wh8, wl8 = cache[key]
        ah = f16(a); al = a - ah
        ah8 = qa_hi(ah * sA) / sA; al8 = qa_lo(al * sA * 2048.0) / (sA * 2048.0)
        return ah @ wh.T + al8 @ wh8.T + ah8 @ wl8.T
    return mm
toks = O.tokenize(seq)[None].repeat(len(pos), 1)
for r, i in enumerate(pos): toks[r, i] = 32
wtid = torch.tensor([O.TOK[seq[i - 1]] for i in pos]); aa = torch.tensor([O.TOK[a] for a in synth.AA20])
def run(dt, rnd=None, mm=None):
    s = O.load_state(st, "esm1v", dt)
    with torch.no_grad():
        lp = torch.log_softmax(O.esm_forward(s, toks, "esm1v", arch.layers, arch.heads, True, dt, rnd, mm), -1)
    rows = torch.stack([lp[r, i] for r, i in enumerate(pos)]).double()
    return rows[:, aa] - rows[torch.arange(len(pos)), wtid][:, None]
ref = run(torch.float64)
cases = [("all four e4m3 (f16f8)", (q8, q8, q8, q8)),
         ("only lo8_A quantised", (q8, ident, ident, ident)),
         ("only hi8_A quantised", (ident, q8, ident, ident)),
         ("only hi8_W quantised", (ident, ident, q8, ident)),
         ("only lo8_W quantised", (ident, ident, ident, q8)),
         ("lo8_W two digits, rest e4m3", (q8, q8, q8, q2)),
         ("W planes two digits (hi8_W, lo8_W)", (q8, q8, q2, q2)),
         ("A planes two digits", (q2, q2, q8, q8)),
         ("cross terms exact, attention split2 (floor)", (ident, ident, ident, ident))]
for name, qs in cases:
    t0 = time.time()
    e = (run(torch.float32, rnd=split2, mm=make_mm(*qs)) - ref).abs()
    print(f"{name:46s} max={e.max().item():.2e} mean={e.mean().item():.2e} ({time.time()-t0:.0f}s)", flush=True)
