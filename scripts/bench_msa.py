"""Throughput of the MSA pre-processing kernels (pg_msa_cluster_neighbors: byte compares/s; pg_msa_prior: GB/s) vs the numpy port."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from proteingym_b200 import _lib
from oracle import tranception_oracle as TO
lib = _lib.load()
for N, L in ((20000, 500), (60000, 300)):
    rng = np.random.RandomState(0)
    base = rng.randint(1, 21, size=(50, L))
    mat = base[rng.randint(0, 50, size=N)]
    mut = rng.rand(N, L) < 0.15
    mat = np.where(mut, rng.randint(0, 21, size=(N, L)), mat).astype(np.uint8)
    d_tok = torch.from_numpy(mat).cuda(); need = torch.full((N,), int(0.8 * L), dtype=torch.int32, device="cuda"); out = torch.empty(N, dtype=torch.int32, device="cuda")
    for _ in range(2): lib.pg_msa_cluster_neighbors(d_tok.data_ptr(), L, N, L, need.data_ptr(), out.data_ptr(), None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    lib.pg_msa_cluster_neighbors(d_tok.data_ptr(), L, N, L, need.data_ptr(), out.data_ptr(), None)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    n_cpu = 300
    t0 = time.time(); TO.cluster_weights(mat[:n_cpu].astype(np.int64), 0.8); t_cpu = (time.time() - t0) * (N / n_cpu) ** 2
    tt = torch.from_numpy(np.ascontiguousarray(mat.T)).cuda(); w = torch.ones(N, dtype=torch.float64, device="cuda"); po = torch.empty(L, 25, dtype=torch.float64, device="cuda")
    for _ in range(2): lib.pg_msa_prior(tt.data_ptr(), w.data_ptr(), N, L, 25, 1e-5, po.data_ptr(), None)
    torch.cuda.synchronize(); e0.record()
    lib.pg_msa_prior(tt.data_ptr(), w.data_ptr(), N, L, 25, 1e-5, po.data_ptr(), None)
    e1.record(); torch.cuda.synchronize(); ms2 = e0.elapsed_time(e1)
    print(json.dumps({"N": N, "L": L, "cluster_ms": round(ms, 2), "byte_compares_per_s": round(N * N * L / ms * 1e3 / 1e12, 2), "unit": "T/s",
                      "numpy_port_extrapolated_s": round(t_cpu, 1), "prior_ms": round(ms2, 3), "prior_GBps": round((N * L + 8 * N * L / 1) / ms2 / 1e6, 1)}))
