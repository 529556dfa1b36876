"""MSA Transformer masked-marginals throughput at MSA-1b size (12 x 768, 12 heads, ffn 3072) with random weights and a synthetic
alignment: ms per masked position, algorithmic TFLOP/s, per-category device time. One JSON line per case.
  python scripts/bench_msa_transformer.py [--rows 400] [--length 512] [--positions 8] [--precision f16f8]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from proteingym_b200 import _lib, checkpoint, msa_engine, synth  # noqa: E402


def flops(arch, R, C):
    d, f, L = arch.embed_dim, arch.ffn_dim, arch.layers
    lin = 2.0 * R * C * (8 * d * d + 2 * d * f) * L            # two attention blocks (4 d^2 each) + FFN
    tied = 2 * 2.0 * C * C * R * d * L                          # scores + context of the tied row attention
    col = 2 * 2.0 * R * R * C * d * L                           # column attention
    return lin, tied, col


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=400)
    ap.add_argument("--length", type=int, default=512)
    ap.add_argument("--positions", type=int, default=8)
    ap.add_argument("--per-pass", type=int, default=4)
    ap.add_argument("--precision", default="f16f8")
    a = ap.parse_args()
    arch = synth.MSA_1B
    cfg = checkpoint.config_from_msa_synth(arch)
    st = checkpoint.normalise_msa_synth_state(arch, synth.make_msa_state(arch, 0, device="cuda"))
    t = synth.random_protein(a.length, seed=1)
    rows = synth.random_alignment(t, a.rows, seed=2)
    toks = msa_engine.tokenize_alignment(rows)
    R, Cc = toks.shape
    sc = msa_engine.MsaScorer(cfg, st, precision=a.precision, max_rows=msa_engine.default_max_rows(cfg, R, Cc, want=a.per_pass))
    pos = np.linspace(1, Cc - 1, a.positions).astype(np.int32)
    sc.masked_marginal_rows(toks, pos[:a.per_pass])  # warm-up
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.pg_profile_begin()
    t0 = time.time()
    out = sc.masked_marginal_rows(toks, pos)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ncat = len(_lib.PROFILE_CATEGORIES)
    ms = (C.c_float * ncat)(); cnt = (C.c_int32 * ncat)()
    lib.pg_profile_end(ms, cnt, ncat)
    lin, tied, col = flops(arch, R, Cc)
    cats = {n: round(float(ms[i]), 2) for i, n in enumerate(_lib.PROFILE_CATEGORIES) if cnt[i]}
    print(json.dumps({"case": f"MSA-1b R={R} C={Cc} positions={a.positions} per_pass={a.per_pass}", "precision": a.precision,
                      "ms_per_position": 1e3 * dt / a.positions, "positions_per_s": a.positions / dt,
                      "algorithmic_tflop_per_position": (lin + tied + col) / 1e12, "share_linear_tied_column": [lin / (lin + tied + col), tied / (lin + tied + col), col / (lin + tied + col)],
                      "algorithmic_tflops": (lin + tied + col) * a.positions / dt / 1e12, "kernel_ms": cats,
                      "finite": bool(torch.isfinite(out).all())}))
    sc.close()


if __name__ == "__main__":
    main()
