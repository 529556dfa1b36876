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
