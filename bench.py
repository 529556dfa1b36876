#!/usr/bin/env python
"""bench.py — mutants/sec of ESM-1v 650M masked-marginal scoring (BASELINE.json metric, config 2).

A "step" is one pass of the hot path over one synthetic assay: WT of L=512 residues, 5000 single mutants
(SURVEY.md §8d config 2): 512 masked copies x 514 tokens through 33 layers, masked-row LM head, mutant scoring.

  python bench.py --gpus N --steps K --warmup W            # our arm  (one JSON line on rank 0)
  python bench.py --impl reference --gpus N ...            # reference arm: the CPU port of the reference path

`value`  : whole-job mutants/s with inputs already resident in HBM (device-timed, max over ranks).
`e2e`    : same metric through the public API (EsmScorer.score_assay) with host buffers: tokenise/parse on host, pinned
           H2D of the int32 block, D2H of the scores, all inside the timed region.
`roofline`: tensor-pipe bound; achieved = algorithmic GEMM FLOPs (2*M*N*K per launch; x1, never the x3 split work)
           / CUDA-event time of the tcgen05 GEMM launches inside the timed region, vs MEASURED_PEAKS.json bf16 sustained.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

L_SEQ, N_MUT, N_ASSAYS = 512, 5000, 10
PASSES = {"f16x3": 3, "f16f8": 2, "f16": 1}  # tensor-pipe units per algorithmic FLOP of a linear layer
METRIC = "mutants/sec (ESM-1v 650M masked-marginal, L<=1024)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f16f8", choices=["f16f8", "f16x3", "f16"],
                    help="f16f8 (headline) and f16x3 = parity modes (<=1e-3 abs vs the fp32 reference); f16 = single-pass fast mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--small", action="store_true", help="tiny model/assay (debugging only; not a valid bench)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), "measured"
    except Exception:
        return 1400.0, 1590.0, "fallback"


def algorithmic_flops(arch, T, P):
    """F_fwd(T) = Lyr*[2*T*(4d^2+2df) + 4*T^2*d] + 2*T*(d^2+d*V), per assay P*F_fwd."""
    Lyr, d, f, V = arch.layers, arch.embed_dim, arch.ffn_dim, arch.vocab
    lin = Lyr * 2 * T * (4 * d * d + 2 * d * f)
    att = Lyr * 4 * T * T * d
    head = 2 * T * (d * d + d * V)
    return P * (lin + att + head), P * lin


def secondary(cats, arch, T, P, steps, precision):
    """Rooflines of the non-dominant kernels from the same event timings: LayerNorm against measured HBM bandwidth
    (algorithmic bytes: read 4d, write 2d per operand plane, per row), attention as algorithmic TFLOP/s (4*T^2*d per layer)."""
    out = {}
    npl = 1 if precision == "f16" else 2
    rows = P * T
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    if "layernorm" in cats and cats["layernorm"]["ms"] > 0:
        byts = 2 * arch.layers * rows * (4 * arch.embed_dim + 2 * arch.embed_dim * npl) * steps
        gbs = byts / (cats["layernorm"]["ms"] / 1e3) / 1e9
        out["layernorm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}
    if "attention" in cats and cats["attention"]["ms"] > 0:
        fl = arch.layers * P * 4.0 * T * T * arch.embed_dim * steps
        out["attention"] = {"bound": "tensor+mufu", "achieved": fl / (cats["attention"]["ms"] / 1e3) / 1e12, "unit": "TFLOP/s (algorithmic)"}
    return out


def make_assay(i, L, n_mut):
    from proteingym_b200 import synth
    seq = synth.random_protein(L, seed=i)
    return seq, synth.sample_mutants(seq, n_mut, seed=1000 + i)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


_CPU_CACHE = {}


def cpu_port_mutants_per_s(arch, state, seconds, threads):
    """Time the CPU port of the reference loop (oracle/esm_oracle.py: batch-1 masked forwards, fp32, all host threads) on a
    bounded sample of config 2 and extrapolate: mutants/s = N_MUT / ((L+2) * t_forward) — the reference runs L+2 forwards
    per checkpoint (compute_fitness.py:489) and forward time does not depend on which position is masked."""
    from oracle import esm_oracle as O
    kind = "esm2" if arch.kind == "esm2" else "esm1v"
    if "st" not in _CPU_CACHE:
        _CPU_CACHE["st"] = O.load_state(state, kind, torch.float32)
    st = _CPU_CACHE["st"]
    # "all the host threads it can use": batch-1 forwards stop scaling (and regress) well before 128 threads, so pick the
    # thread count that is fastest on a short probe instead of handicapping the CPU arm with oversubscription
    probe = O.tokenize(make_assay(0, 128 if arch.layers > 8 else 32, 10)[0])[None]
    best = _CPU_CACHE.get("best", (None, float("inf")))
    for nt in ([] if best[0] else sorted({t for t in (8, 16, 32, 64) if t <= threads} or {threads})):
        torch.set_num_threads(nt)
        with torch.no_grad():
            O.esm_forward(st, probe, kind, arch.layers, arch.heads, arch.token_dropout)
            t0 = time.time()
            O.esm_forward(st, probe, kind, arch.layers, arch.heads, arch.token_dropout)
            dt = time.time() - t0
        if dt < best[1]:
            best = (nt, dt)
    _CPU_CACHE["best"] = best
    threads = best[0]
    torch.set_num_threads(threads)
    log(f"cpu port: {threads} threads (probe {best[1]:.3f} s)")
    seq, _ = make_assay(0, L_SEQ if arch.layers > 8 else 64, 10)
    toks = O.tokenize(seq)[None]
    times = []
    t_start = time.time()
    i = 0
    with torch.no_grad():
        while True:
            tb = toks.clone(); tb[0, 1 + i] = O.MASK_IDX
            t0 = time.time()
            lp = torch.log_softmax(O.esm_forward(st, tb, kind, arch.layers, arch.heads, arch.token_dropout), -1)[:, 1 + i]
            times.append(time.time() - t0)
            i += 1
            if i >= 3 and time.time() - t_start > seconds:
                break
            if i >= 64:
                break
    t_fwd = statistics.median(times[1:]) if len(times) > 1 else times[0]
    T = toks.shape[1]
    n_mut = N_MUT if arch.layers > 8 else 200
    return n_mut / (T * t_fwd), {"forwards_timed": len(times), "t_forward_s": t_fwd, "T": T, "threads": threads}


def main():
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    a = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from proteingym_b200 import synth
    arch = synth.EsmArch("esm1v", 2, 128, 2, 256) if a.small else synth.ESM1V_650M
    L = 64 if a.small else L_SEQ
    n_mut = 200 if a.small else N_MUT
    T = L + 2
    config = {"workload": f"config2: {N_ASSAYS} synthetic assays, L={L} (T={T}) x {n_mut} single mutants, ESM-1v 650M "
                          f"({arch.layers}x{arch.embed_dim}, {arch.heads} heads, ffn {arch.ffn_dim}), masked-marginals, "
                          "one assay per step",
              "assays_per_step": 1, "mutants_per_step": n_mut, "masked_positions_per_step": L,
              "parallelism": f"assay-sharded x{a.gpus} (weights broadcast once, scores gathered once)",
              "l2": "no flush needed: per-step working set (fp16 weights 1.3-3.9 GB + activations >6 GB) >> 126 MB L2"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        state = synth.make_esm_state(arch, seed=0)
        vals, info = [], None
        per_step = max(2.0, min(a.cpu_seconds, 150.0 / max(1, a.steps + a.warmup)))
        for s in range(a.warmup + a.steps):
            v, info = cpu_port_mutants_per_s(arch, state, per_step, threads)
            if s >= a.warmup:
                vals.append(v)
        v = statistics.median(vals)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "mutants/s", "n_gpus": a.gpus,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * n_mut / v, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "mutants/s", "cores": info["threads"], "kind": "port",
                                           "sample": f"{info['forwards_timed']} batch-1 masked forwards of T={info['T']} per step "
                                                     f"(median {info['t_forward_s']:.3f} s) extrapolated to the reference's "
                                                     f"L+2={info['T']} forwards per assay"},
                          "e2e": {"value": v, "unit": "mutants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from proteingym_b200 import _lib, checkpoint
    from proteingym_b200.esm_engine import EsmScorer
    lib = _lib.load()

    # weights: rank 0 builds the seeded synthetic checkpoint, NCCL-broadcasts it (north_star: "broadcast of weights")
    t_w0 = time.time()
    state = None
    log("building synthetic weights")
    if rank == 0:
        state = checkpoint.normalise_synth_state(arch, synth.make_esm_state(arch, seed=0))
    if world > 1:
        from proteingym_b200 import sharding
        state = sharding.broadcast_state(state, src=0, device=torch.device("cuda", local_rank))
    sustained, burst, how = peaks()
    total = a.warmup + a.steps
    my_assays = [make_assay((rank + world * s) % N_ASSAYS, L, n_mut) for s in range(total)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(precision, with_e2e):
        """One full measurement (device-resident leg, optional end-to-end leg) at the given operand precision."""
        log(f"measure {precision}: upload")
        t_w0 = time.time()
        scorer = EsmScorer(checkpoint.config_from_synth(arch), state, precision=precision, device=local_rank,
                           max_rows=16384 if a.small else 0)
        load_s = time.time() - t_w0
        preps = [scorer.prepare_assay(seq, muts) for seq, muts in my_assays]
        devs = [h.to(scorer.device) for h, _ in preps]
        sampler = ClockSampler(local_rank) if rank == 0 else None
        # ---- leg 1: HBM-resident (value) + per-kernel event timing for the roofline ----
        log(f"measure {precision}: resident leg")
        for s in range(a.warmup):
            scorer.run_assay(preps[s][0], preps[s][1], dev=devs[s])
        barrier()
        launches0 = lib.pg_launch_count()
        lib.pg_profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.time()
        e0.record()
        outs = []
        for s in range(a.warmup, total):
            outs.append(scorer.run_assay(preps[s][0], preps[s][1], dev=devs[s]))
        if dist is not None:  # the single final gather of the per-mutant scores
            mine = torch.stack(outs)
            gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
            dist.gather(mine, gathered, dst=0)
        e1.record()
        barrier()
        t_host1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        launches = lib.pg_launch_count() - launches0
        ncat = len(_lib.PROFILE_CATEGORIES)
        cat_ms = (C.c_float * ncat)(); cat_n = (C.c_int32 * ncat)()
        lib.pg_profile_end(cat_ms, cat_n, ncat)
        clocks = sampler.summary(t_host0, t_host1) if sampler else None
        res = {"value": world * a.steps * n_mut / (ms_total / 1e3), "ms_per_step": ms_total / a.steps, "gpu_launches": int(launches),
               "clocks": clocks, "weight_load_s": load_s}
        # ---- leg 2: end to end through the public API with host buffers ----
        if with_e2e:
            log(f"measure {precision}: e2e leg")
            for s in range(min(2, a.warmup)):
                scorer.score_assay(*my_assays[s])
            barrier()
            t0 = time.time()
            e2e_scores = []
            for s in range(a.warmup, total):
                e2e_scores.append(scorer.score_assay(*my_assays[s]))
            barrier()
            e2e_s = torch.tensor([time.time() - t0], device="cuda")
            if dist is not None:
                dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
            same = all(np.array_equal(e2e_scores[i], outs[i].cpu().numpy()) for i in range(len(outs)))
            res["e2e"] = {"value": world * a.steps * n_mut / float(e2e_s.item()), "unit": "mutants/s",
                          "h2d_bytes_per_step": int(preps[a.warmup][0].numel() * 4), "d2h_bytes_per_step": int(n_mut * 4),
                          "bit_identical_to_resident_leg": bool(same)}
        # ---- roofline of the dominant kernel (tcgen05 GEMM) from the event timings inside the timed region ----
        cats = {n: {"ms": float(cat_ms[i]), "launches": int(cat_n[i])} for i, n in enumerate(_lib.PROFILE_CATEGORIES) if cat_n[i]}
        P = preps[a.warmup][1]["P"]
        f_total, f_lin = algorithmic_flops(arch, T, P)
        gemm_ms = sum(cats[c]["ms"] for c in cats if c.startswith("gemm_"))
        gemm_launches = sum(cats[c]["launches"] for c in cats if c.startswith("gemm_"))
        achieved = (f_lin * a.steps / 1e12) / (gemm_ms / 1e3) if gemm_ms > 0 else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                traffic = json.load(fh).get(precision)
        except Exception:
            pass
        passes = PASSES[precision]
        res["roofline"] = {
            "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05.mma kind::f16, M128xN256xK16, TMA 4-stage, TMA-store epilogue)",
            "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": (achieved / sustained) if achieved else None,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({how}); burst {burst}",
            "algorithmic_flops_per_launch": f_lin * a.steps / max(1, gemm_launches), "launches": gemm_launches,
            "avg_launch_ms": gemm_ms / max(1, gemm_launches), "tensor_pipe_work_multiplier": passes,
            "issued_tflops": (achieved * passes) if achieved else None,
            "issued_frac": (achieved * passes / sustained) if achieved else None, "traffic": traffic,
            "whole_step": {"algorithmic_tflop_per_step": f_total / 1e12,
                           "achieved_per_gpu": f_total * a.steps / 1e12 / (ms_total / 1e3),
                           "frac_of_peak": f_total * a.steps / 1e12 / (ms_total / 1e3) / sustained},
            "kernel_ms_in_timed_region": cats,
            "secondary": secondary(cats, arch, T, P, a.steps, precision)}
        scorer.close()
        del scorer, devs
        torch.cuda.empty_cache()
        return res

    cpu_state = {k: v.cpu() for k, v in state.items()} if (rank == 0 and world == 1 and not a.no_cpu_baseline) else None
    main_res = measure(a.precision, with_e2e=True)
    others = [] if a.small else [(m, measure(m, with_e2e=False)) for m in ("f16f8", "f16x3", "f16") if m != a.precision]
    del state

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    DT = {"f16f8": "f16f8 (linear layers: fp16 hi*hi + e4m3 cross terms = 2 tcgen05 passes-equivalents; attention fp16 hi/lo x3; chunked "
                   "fp32 accumulation; fp32 residual/softmax; meets 1e-3 parity)",
          "f16x3": "f16x3 (fp16 hi+lo operand pairs, 3 tcgen05 passes, fp32 accumulate/residual/softmax; meets 1e-3 parity)",
          "f16": "f16 (single fp16 pass, fp32 accumulate; ~1e-2 abs error, Spearman > 0.999; does NOT meet the 1e-3 parity bar)"}
    out = {"metric": METRIC, "value": main_res["value"], "unit": "mutants/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": DT[a.precision], "data": "synthetic", "config": config, "precision_mode": a.precision,
           "e2e": main_res["e2e"], "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"],
           "roofline": main_res["roofline"], "weight_load_s": main_res["weight_load_s"]}
    out["other_precision_modes"] = [
        {"precision_mode": m, "dtype": DT[m], "value": r["value"], "unit": "mutants/s", "ms_per_step": r["ms_per_step"], "clocks": r["clocks"],
         "roofline": {k: r["roofline"][k] for k in ("achieved", "frac", "issued_tflops", "issued_frac", "whole_step",
                                                    "kernel_ms_in_timed_region", "secondary")}} for m, r in others]

    if not a.no_cpu_baseline and world == 1:
        log("cpu baseline")
        threads = os.cpu_count() or 1
        v, info = cpu_port_mutants_per_s(arch, cpu_state, a.cpu_seconds, threads)
        out["cpu_baseline"] = {"value": v, "unit": "mutants/s", "cores": info["threads"], "host_cpus": threads, "kind": "port",
                               "sample": f"{info['forwards_timed']} batch-1 masked forwards of T={info['T']} (median "
                                         f"{info['t_forward_s']:.3f} s) extrapolated to the reference's L+2 forwards per assay"}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
