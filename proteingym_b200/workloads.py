"""Synthetic workloads shaped like ProteinGym's benchmarks (BASELINE.json configs 3-5), for bench.py and the multi-assay tests.

Shapes (sequence length, number of mutants, number of multi-mutants of every DMS_substitutions / DMS_indels assay) come from
``tests/golden/dms_workload_shapes.json`` (written by oracle/gen_workload_shapes.py from the reference's reference_files);
wild types are random proteins of those lengths and mutant lists are drawn to size — no network, no real data.
A bench cannot run all 217 assays inside its time budget, so ``pick_subset`` takes a bounded, deterministic sample and says
what it left out."""
from __future__ import annotations

import json
import os

import numpy as np

from . import sharding, synth

SHAPES_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dms_workload_shapes.json")


def load_shapes(path: str = SHAPES_PATH) -> dict:
    with open(path) as fh:
        return json.load(fh)


def esm_cost(entry: dict, arch) -> float:
    """Algorithmic FLOPs of one masked-marginal assay: P = L masked positions of T = min(L + 2, 1024) tokens."""
    return sharding.assay_cost(entry["L"], arch.layers, arch.embed_dim, arch.ffn_dim)


def tranception_cost(entry: dict, arch, n_mutants: int) -> float:
    """Linear + causal-attention FLOPs of scoring ``n_mutants`` sequences (+ 1 WT) in both directions, T = min(L, n_ctx - 2) + 2."""
    T = min(entry["L"], arch.n_ctx - 2) + 2
    d, f = arch.embed_dim, arch.ffn_dim
    per_tok = arch.layers * (2.0 * (4 * d * d + 2 * d * f) + 2.0 * T * d)
    return 2.0 * (n_mutants + 1) * T * per_tok


def pick_subset(entries, costs, k: int, cost_cap: float, replicas: int = 1):
    """k size classes spread evenly over the cost-sorted list of the entries with cost <= cost_cap, ``replicas`` neighbouring entries
    of that list per class (weak scaling: N ranks score N distinct assays of each class, so an LPT assignment of the subset is as
    balanced as one of the full benchmark; replicas = 1 is the plain evenly-spaced pick). Returns (indices, info)."""
    order = [i for i in sorted(range(len(entries)), key=lambda i: (costs[i], i)) if costs[i] <= cost_cap]
    dropped = len(entries) - len(order)
    k = min(k, len(order))
    if k == 0:
        return [], {"eligible": 0, "dropped_over_cap": dropped}
    n = len(order)
    centres = [int(round(j * (n - 1) / max(1, k - 1))) for j in range(k)] if k > 1 else [n // 2]
    picks = set()
    for c in centres:
        start = min(max(c - replicas // 2, 0), max(n - replicas, 0))
        picks.update(order[start:start + replicas])
    picks = sorted(picks)
    info = {"eligible": n, "dropped_over_cap": dropped, "cost_cap_tflop": cost_cap / 1e12, "classes": k, "replicas": replicas,
            "subset_cost_tflop": sum(costs[i] for i in picks) / 1e12, "all_cost_tflop": sum(costs) / 1e12}
    return picks, info


def substitution_assay(entry: dict, seed: int, max_mutants: int):
    """(wild type, mutant strings): uniform singles, with the assay's share of 2-5-site multi-mutants when it has any."""
    seq = synth.random_protein(entry["L"], seed=seed)
    n = min(entry["n_mutants"], max_mutants, 19 * entry["L"])
    frac = entry["n_multi"] / max(1, entry["n_mutants"])
    return seq, synth.sample_mutants(seq, n, seed=1000 + seed, multi_frac=frac)


def indel_assay(entry: dict, seed: int, max_mutants: int):
    seq = synth.random_protein(entry["L"], seed=seed)
    return seq, synth.random_indels(seq, min(entry["n_mutants"], max_mutants), seed=1000 + seed)


def synthetic_log_prior(L: int, seed: int, vocab: int = 25) -> np.ndarray:
    """[L, vocab] fp32 log-probabilities over the amino-acid columns (>= 5), -inf on the special tokens — the shape of the
    retrieval priors get_msa_prior / the EVE decoder produce (trancepteve/model_pytorch.py:940-1001)."""
    rng = np.random.RandomState(seed)
    p = rng.dirichlet(np.ones(vocab - 5), size=L)
    out = np.full((L, vocab), -np.inf, dtype=np.float32)
    out[:, 5:] = np.log(p)
    return out
