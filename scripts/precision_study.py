"""Numerics study (CPU): emulate tensor-core operand rounding on the oracle at true ESM-1v size and report the
per-mutant score error vs an fp64 run. Decides the GEMM operand format (fp16 vs bf16 vs split) against the 1e-3 target."""
import sys, time
sys.path.insert(0, "/root/repo")
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
f16 = lambda t: t.to(torch.float16).to(t.dtype)
bf16 = lambda t: t.to(torch.bfloat16).to(t.dtype)
def split2(t):  # hi+lo fp16 pair == ~22-bit operand
    hi = t.to(torch.float16).to(t.dtype); lo = (t - hi).to(torch.float16).to(t.dtype); return hi + lo
res = {}
for name, dt, rnd in (("fp64", torch.float64, None), ("fp32", torch.float32, None), ("fp16", torch.float32, f16),
                      ("bf16", torch.float32, bf16), ("fp16x2", torch.float32, split2)):
    t = time.time()
    s = O.load_state(st, "esm1v", dt)
    toks = O.tokenize(seq)[None].repeat(len(pos), 1)
    for r, i in enumerate(pos): toks[r, i] = 32
    with torch.no_grad():
        lp = torch.log_softmax(O.esm_forward(s, toks, "esm1v", arch.layers, arch.heads, True, dt, rnd), -1)
    rows = torch.stack([lp[r, i] for r, i in enumerate(pos)]).double()
    res[name] = rows
    print(name, "%.1fs" % (time.time() - t), flush=True)
ref = res["fp64"]
wt = torch.tensor([O.TOK[seq[i - 1]] for i in pos])
aa = torch.tensor([O.TOK[a] for a in synth.AA20])
def scores(rows): return rows[:, aa] - rows[torch.arange(len(pos)), wt][:, None]
sr = scores(ref)
print("score range", sr.min().item(), sr.max().item(), "std", sr.std().item())
for k, v in res.items():
    if k == "fp64": continue
    e = (scores(v) - sr).abs()
    print(f"{k:7s} max|dscore|={e.max().item():.2e} mean={e.mean().item():.2e}  max|dlogp|={(v-ref)[:, aa].abs().max().item():.2e}")
