"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY. Build-container script (needs /root/reference).
Golden vectors for the MSA pre-processing rows from the UNMODIFIED reference functions:
  * tranception/utils/msa_utils.py::get_msa_prior on a synthetic a2m (unweighted and with a name->weight subset is not
    reachable without MSA_processing, so only the unweighted call is pinned);
  * proteingym/utils/weights.py::calc_weights_fast / calc_num_cluster_members_nogaps executed as plain Python with a stub
    ``numba`` module (numba is not installed here; @jit -> identity, prange -> range)."""
import importlib.util
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import ref_shims_tranception as R  # noqa: E402
from proteingym_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    mp, su = R.install()
    from tranception.utils import msa_utils
    seq = synth.random_protein(70, 41)
    msa = synth.synthetic_msa(seq, 60, seed=2)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "msa.a2m")
    synth.write_a2m(path, msa)
    vocab = R.tokenizer().get_vocab()
    prior = msa_utils.get_msa_prior(path, None, 5, 75, 90, vocab, retrieval_aggregation_mode="aggregate_substitution", filter_MSA=True)
    np.save(os.path.join(GOLD, "msa_prior_reference.npy"), prior)
    # sequence weights: run the reference source with a stub numba
    nb = types.ModuleType("numba")
    nb.jit = lambda *a, **k: (lambda f: f)
    nb.prange = range
    nb.set_num_threads = lambda n: None
    nb.get_num_threads = lambda: 1
    nb.config = types.SimpleNamespace(NUMBA_NUM_THREADS=1)
    sys.modules["numba"] = nb
    spec = importlib.util.spec_from_file_location("pg_ref_weights", os.path.join(R.REF, "..", "..", "utils", "weights.py"))
    W = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(W)
    W.range = lambda *a: range(*[int(x) for x in a])  # numba accepts the float L of `L = 1.0 * L; range(L)`; CPython does not
    rng = np.random.RandomState(3)
    base = rng.randint(1, 21, size=(6, 40))
    rates = (0.02, 0.05, 0.1, 0.15, 0.2, 0.3)
    mat = np.concatenate([np.where(rng.rand(12, 40) < rates[k], rng.randint(1, 21, size=(12, 40)), base[k]) for k in range(6)])
    mat[rng.rand(*mat.shape) < 0.1] = 0
    mat[5] = 0  # an all-gap sequence -> weight 0
    w = W.calc_weights_fast(mat.astype(np.int64), identity_threshold=0.8, empty_value=0, num_cpus=1)
    np.savez(os.path.join(GOLD, "msa_weights_reference.npz"), matrix=mat.astype(np.int8), weights=w)
    json.dump({"target_seq": seq, "msa_seed": 2, "msa_n": 60, "MSA_start": 5, "MSA_end": 75, "len_target_seq": 90, "identity_threshold": 0.8},
              open(os.path.join(GOLD, "msa_meta.json"), "w"), indent=1)
    print("[gen_golden_msa] prior", prior.shape, "weights", w[:8])


if __name__ == "__main__":
    main()
