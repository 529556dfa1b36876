"""Tranception-L (36 x 1280, 20 heads, n_ctx 1024) autoregressive scoring throughput on one B200 (BASELINE config 4 shape:
synthetic assays, random-init weights). mutants/s counts both scoring directions and the WT windows, through
TranceptionScorer.score_mutants (host slicing + pg_ar_loglik). Prints one JSON line per case."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from proteingym_b200 import _lib, synth
from proteingym_b200.tranception_engine import TranceptionScorer

arch = synth.TRANCEPTION_L
st = {k[len("transformer."):]: v for k, v in synth.make_tranception_state(arch, 0).items() if k.startswith("transformer.")}
cfg = {"n_embd": arch.embed_dim, "n_head": arch.heads, "n_layer": arch.layers, "n_ctx": arch.n_ctx, "n_inner": arch.ffn_dim, "vocab_size": 25}
lib = _lib.load()
cases = [("subs_L512_1000", 512, 1000, False), ("indels_L62_2000", 62, 2000, True), ("subs_L1500_windowed_300", 1500, 300, False)]
if os.environ.get("PG_BENCH_CASES"): cases = [c for c in cases if c[0] in os.environ["PG_BENCH_CASES"].split(",")]
PRECS = os.environ.get("PG_BENCH_PRECS", "f16x3,f16").split(",")
for precision in PRECS:
    sc = TranceptionScorer(cfg, st, precision=precision)
    for name, L, n, indel in cases:
        seq = synth.random_protein(L, 7)
        if indel:
            v = synth.random_indels(seq, n, 3)
            dms = pd.DataFrame({"mutant": v, "mutated_sequence": v})
        else:
            m = synth.sample_mutants(seq, n, 3)
            dms = pd.DataFrame({"mutant": m, "mutated_sequence": [synth.apply_mutant(seq, x) for x in m]})
        sc.score_mutants(dms.iloc[:50], seq, indel_mode=indel)  # warm-up
        torch.cuda.synchronize()
        ncat = len(_lib.PROFILE_CATEGORIES); ms = (C.c_float * ncat)(); cnt = (C.c_int32 * ncat)()
        lib.pg_profile_begin(); t0 = time.time()
        out = sc.score_mutants(dms, seq, indel_mode=indel)
        torch.cuda.synchronize(); dt = time.time() - t0
        lib.pg_profile_end(ms, cnt, ncat)
        T = min(L, 1022) + 2
        tokens = 2 * (n + (1 if L <= 1022 else n)) * T  # both directions, + WT windows (upper bound when windowed)
        flops = tokens * (arch.layers * (2 * (4 * arch.embed_dim ** 2 + 2 * arch.embed_dim * arch.ffn_dim) + 2 * T * arch.embed_dim))
        print(json.dumps({"case": name, "precision": precision, "mutants": n, "seconds": round(dt, 3), "mutants_per_s": round(n / dt, 1),
                          "algorithmic_tflops": round(flops / dt / 1e12, 1),
                          "kernel_ms": {c: round(float(ms[i]), 1) for i, c in enumerate(_lib.PROFILE_CATEGORIES) if cnt[i]}}), flush=True)
    sc.close()
