"""CPU, world_size 2, gloo: the N>1 host logic (weight broadcast, LPT assay assignment, final score gather)."""
import os

import numpy as np
import torch
import torch.multiprocessing as mp

from proteingym_b200 import sharding, synth


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arch = synth.EsmArch("esm1v", 1, 64, 1, 64)
    state = synth.make_esm_state(arch, seed=5) if rank == 0 else None
    if state is not None:
        state.pop("lm_head.weight")
    got = sharding.broadcast_state(state, src=0)
    ref = synth.make_esm_state(arch, seed=5)
    ok_bcast = all(torch.equal(got[k], ref[k]) for k in got) and set(got) == set(ref) - {"lm_head.weight"}
    lens = [37, 512, 90, 1500, 245, 245, 64, 800]
    costs = [sharding.assay_cost(L, 33, 1280, 5120) for L in lens]
    assign = sharding.lpt_assign(costs, world)
    mine = {i: torch.full((10 + i,), float(i)) + torch.arange(10 + i) * 0.5 for i in assign[rank]}
    merged = sharding.gather_scores(mine, dst=0)
    q.put((rank, ok_bcast, assign, None if merged is None else {k: v.tolist() for k, v in merged.items()}))
    dist.destroy_process_group()


def test_broadcast_assign_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    a0, a1 = res[0][2], res[1][2]
    assert a0 == a1  # both ranks computed the same assignment
    assert sorted(a0[0] + a0[1]) == list(range(8)) and not set(a0[0]) & set(a0[1])
    merged = res[0][3]
    assert res[1][3] is None and sorted(merged) == list(range(8))
    for i in range(8):
        assert np.allclose(merged[i], np.full(10 + i, float(i)) + np.arange(10 + i) * 0.5)


def test_lpt_balances_real_length_distribution():
    rng = np.random.RandomState(0)
    lens = np.clip(rng.lognormal(5.5, 0.8, 217).astype(int), 37, 3423)
    costs = [sharding.assay_cost(int(L), 33, 1280, 5120) for L in lens]
    for world in (2, 4, 8):
        assign = sharding.lpt_assign(costs, world)
        loads = [sum(costs[i] for i in a) for a in assign]
        assert sorted(i for a in assign for i in a) == list(range(217))
        assert max(loads) / (sum(loads) / world) < 1.05  # within 5% of perfect balance
    assert sharding.lpt_assign([], 4) == [[], [], [], []]
    assert sharding.lpt_assign([3.0], 2) == [[0], []]


def _worker_rows(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for P in (0, 1, 5, 8, 33):
        lo, hi, chunk = sharding.position_chunk(P, world, rank)
        full = torch.full((max(world * chunk, 1), 3), float("nan"))
        for r in range(lo, hi):
            full[r] = torch.tensor([r, 2.0 * r, -1.0 * r])
        sharding.all_gather_rows(full, chunk)
        want = torch.stack([torch.tensor([r, 2.0 * r, -1.0 * r]) for r in range(P)]) if P else torch.zeros((0, 3))
        ok = ok and torch.equal(full[:P], want)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_position_partition_chunks_and_row_gather_world2():
    """Secondary partitioning of one assay: contiguous position chunks + one all-gather of the log-prob rows."""
    for P, world in ((0, 4), (1, 4), (512, 8), (513, 8), (7, 8), (1022, 3)):
        spans = [sharding.position_chunk(P, world, r) for r in range(world)]
        assert all(c == spans[0][2] for _, _, c in spans)
        covered = [i for lo, hi, _ in spans for i in range(lo, hi)]
        assert covered == list(range(P))                      # disjoint, ordered, complete
        assert all(hi - lo <= spans[0][2] for lo, hi, _ in spans)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_rows, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def _worker_tranception_rows(rank, world, port, q):
    """TranceptEVE scoring with the sequence rows of each direction split over 2 ranks (CPU stand-in for the device half)."""
    import json
    import sys
    import tempfile
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(4)
    from conftest import GOLDEN
    from cpu_trancepteve import CpuTranceptEVE
    from trancepteve_cases import SMALL_CASES as CASES, make_inputs
    name = "trancepteve_long"
    case = CASES[name]
    gd = os.path.join(GOLDEN, name)
    meta = json.load(open(os.path.join(gd, "meta.json")))
    a = case["arch"]
    arch = synth.TranceptionArch(a[0], a[1], a[2], a[3], n_ctx=a[4])
    inp = make_inputs(case, tempfile.mkdtemp())
    sc = CpuTranceptEVE(arch, synth.make_tranception_state(arch, meta["tranception_seed"]), meta["target_seq"], case["kind"],
                        np.load(os.path.join(gd, "msa_log_prior_init.npy")), np.load(os.path.join(gd, "eve_log_prior_init.npy")),
                        case["msa"][0], case["msa"][1], (meta["MSA_processed_depth"], meta["EVE_processed_depth"]), meta["focus_cols"],
                        case["col_thr"], case["msa_recal"], case["eve_recal"])
    sc.device = torch.device("cpu")
    sc.shard = (rank, world)
    out = sc.score_mutants(inp["dms"], meta["target_seq"])
    import pandas as pd
    ref = pd.read_csv(os.path.join(gd, "reference_scores.csv"))
    err = float(np.abs(out["avg_score"].values - ref["avg_score"].values).max())
    q.put((rank, err, list(out["mutated_sequence"]) == list(ref["mutated_sequence"]), out["avg_score"].values.tolist()))
    dist.destroy_process_group()


def test_tranception_rows_split_over_two_ranks_match_reference():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_tranception_rows, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] < 2e-5 and res[1][1] < 2e-5 and res[0][2] and res[1][2]
    assert res[0][3] == res[1][3]   # every rank ends up with the same, complete score vector
