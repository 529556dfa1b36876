import sys, time
sys.path.insert(0, "/root/repo")
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
    hi = t.to(torch.float16).to(t.dtype); lo = (t - hi).to(torch.float16).to(t.dtype); return hi + lo
def e5m2(t): return t.to(torch.float8_e5m2).to(t.dtype)
def hi_plus_lo8(t):  # fp16 hi + e5m2 lo
    hi = f16(t); return hi + e5m2(t - hi)
# weights are identified by being 2-D parameters passed through rnd(W(...)): tag by shape heuristic (ndim==2 and no batch)
def mk(act, wt):
    def rnd(t):
        return wt(t) if t.ndim == 2 else act(t)
    return rnd
toks = O.tokenize(seq)[None].repeat(len(pos), 1)
for r, i in enumerate(pos): toks[r, i] = 32
wtid = torch.tensor([O.TOK[seq[i - 1]] for i in pos]); aa = torch.tensor([O.TOK[a] for a in synth.AA20])
def run(dt, rnd):
    s = O.load_state(st, "esm1v", dt)
    with torch.no_grad():
        lp = torch.log_softmax(O.esm_forward(s, toks, "esm1v", arch.layers, arch.heads, True, dt, rnd), -1)
    rows = torch.stack([lp[r, i] for r, i in enumerate(pos)]).double()
    return rows[:, aa] - rows[torch.arange(len(pos)), wtid][:, None]
ref = run(torch.float64, None)
for name, rnd in (("act16,w16", mk(f16, f16)), ("act_split,w16", mk(split2, f16)), ("act16,w_split", mk(f16, split2)),
                  ("hi+lo8 both", mk(hi_plus_lo8, hi_plus_lo8))):
    e = (run(torch.float32, rnd) - ref).abs()
    print(f"{name:16s} max={e.max().item():.2e} mean={e.mean().item():.2e}", flush=True)
print("--- attention-operand ablation (GEMMs split) ---")
def mk3(w, a3, a4):
    def rnd(t):
        return w(t) if t.ndim == 2 else (a3(t) if t.ndim == 3 else a4(t))
    return rnd
bf16 = lambda t: t.to(torch.bfloat16).to(t.dtype)
ident = lambda t: t
for name, rnd in (("gemm split, attn f16", mk3(split2, split2, f16)), ("gemm split, attn exact", mk3(split2, split2, ident)),
                  ("gemm exact, attn f16", mk3(ident, ident, f16))):
    e = (run(torch.float32, rnd) - ref).abs()
    print(f"{name:24s} max={e.max().item():.2e} mean={e.mean().item():.2e}", flush=True)
print("--- which attention operands (GEMMs exact) ---")
def mk4(fq, fk, fv, fp):
    cnt = [0]
    def rnd(t):
        if t.ndim != 4: return t
        f = (fq, fk, fv, fp)[cnt[0] % 4]; cnt[0] += 1
        return f(t)
    return rnd
for name, rnd in (("q,k f16 only", mk4(f16, f16, ident, ident)), ("v f16 only", mk4(ident, ident, f16, ident)),
                  ("p f16 only", mk4(ident, ident, ident, f16)), ("k f16 only", mk4(ident, f16, ident, ident))):
    e = (run(torch.float32, rnd) - ref).abs()
    print(f"{name:24s} max={e.max().item():.2e} mean={e.mean().item():.2e}", flush=True)
