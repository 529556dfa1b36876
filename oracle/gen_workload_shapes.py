"""ORACLE SUPPORT — build-container script (needs /root/reference). Extracts the SHAPES of ProteinGym's assays — sequence length,
number of mutants, number of multi-mutants — from reference_files/DMS_substitutions.csv and DMS_indels.csv into
tests/golden/dms_workload_shapes.json. bench.py sizes its synthetic config-3/4/5 workloads (BASELINE.json) from this table;
no sequence or score data is copied (wild types are drawn at random with the listed lengths)."""
import json
import os
import sys

import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PG_REFERENCE_ROOT", "/root/reference")


def main():
    subs = pd.read_csv(os.path.join(REF, "reference_files", "DMS_substitutions.csv"))
    ind = pd.read_csv(os.path.join(REF, "reference_files", "DMS_indels.csv"))
    out = {"source": "reference_files/DMS_substitutions.csv, DMS_indels.csv (shapes only)",
           "substitutions": [{"id": r.DMS_id, "L": int(r.seq_len), "n_mutants": int(r.DMS_total_number_mutants),
                              "n_multi": int(r.DMS_number_multiple_mutants) if r.includes_multiple_mutants else 0}
                             for r in subs.itertuples()],
           "indels": [{"id": r.DMS_id, "L": int(r.seq_len), "n_mutants": int(r.DMS_total_number_mutants)} for r in ind.itertuples()]}
    path = os.path.join(ROOT, "tests", "golden", "dms_workload_shapes.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(path, len(out["substitutions"]), len(out["indels"]), os.path.getsize(path))


if __name__ == "__main__":
    sys.exit(main())
