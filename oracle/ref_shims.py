"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).

Makes the UNMODIFIED reference ESM scorer importable/runnable on CPU with torch 2.11:
  1. stub ``Bio`` modules (Biopython is absent; imported at compute_fitness.py:9 and utils/msa_utils.py:9-11); ``SeqIO.parse`` gets a
     minimal FASTA reader (records with ``.description`` / ``.seq``), which is all the MSA Transformer path asks of it
     (compute_fitness.py:31-37),
  2. ``torch.Tensor.cuda`` -> identity, because masked-marginals calls ``.cuda()`` unconditionally
     (compute_fitness.py:502) even with ``--nogpu``,
  3. ``torch.serialization.add_safe_globals([argparse.Namespace])`` (torch>=2.6 weights_only default vs
     esm/pretrained.py:70),
  4. ``sys.path`` as if the script had been launched from its own directory (compute_fitness.py:14-16).
Nothing from the reference is copied; it is imported from where it lies.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

REF = os.environ.get("PG_REFERENCE_ROOT", "/root/reference")
ESM_DIR = os.path.join(REF, "proteingym", "baselines", "esm")


def available() -> bool:
    return os.path.isfile(os.path.join(ESM_DIR, "compute_fitness.py"))


class _Record:
    def __init__(self, description, seq):
        self.description, self.id, self.seq = description, description.split()[0] if description else "", seq


def _fasta_records(filename, fmt="fasta"):
    """Stand-in for Bio.SeqIO.parse(filename, "fasta"): yields records in file order."""
    assert fmt == "fasta"
    name, parts = None, []
    with open(filename) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    yield _Record(name, "".join(parts))
                name, parts = line[1:], []
            elif name is not None:
                parts.append(line.strip())
    if name is not None:
        yield _Record(name, "".join(parts))


def install():
    """Idempotent. Returns the reference ``compute_fitness`` module."""
    import torch
    if "pg_ref_compute_fitness" in sys.modules:
        return sys.modules["pg_ref_compute_fitness"]
    for name in ("Bio", "Bio.SeqIO", "Bio.SeqRecord", "Bio.Seq", "Bio.Align", "Bio.Align.Applications"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["Bio.SeqIO"].parse = _fasta_records
    sys.modules["Bio.SeqRecord"].SeqRecord = object
    sys.modules["Bio.Seq"].Seq = object
    torch.serialization.add_safe_globals([argparse.Namespace])
    torch.Tensor.cuda = lambda self, *a, **k: self
    for p in (ESM_DIR, os.path.join(REF, "proteingym"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib.util
    spec = importlib.util.spec_from_file_location("pg_ref_compute_fitness", os.path.join(ESM_DIR, "compute_fitness.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pg_ref_compute_fitness"] = mod
    spec.loader.exec_module(mod)
    return mod


def run_reference_cli(argv):
    """Run the reference's ``main`` with its own argparse on ``argv`` (list of strings)."""
    mod = install()
    args = mod.create_parser().parse_args(argv)
    mod.main(args)
