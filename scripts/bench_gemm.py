"""Isolated timing of the four linear layers of an ESM-1v 650M pass (M = 255 x 514 rows) in every operand mode and for several
accumulation-chunk lengths: CUDA events, 3 warm-up + 10 timed launches, L2 flushed between launches by the working set itself
(A + out > 126 MB). Prints one JSON line per (shape, mode, kchunk). Usage: python scripts/bench_gemm.py [kchunk,kchunk,...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from proteingym_b200 import _lib
lib = _lib.load()
M = 255 * 514
SHAPES = [("qkv", 3840, 1280, 0), ("out", 1280, 1280, 2), ("fc1", 5120, 1280, 1), ("fc2", 1280, 5120, 2)]
KCH = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024]
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
for name, N, K, epi in SHAPES:
    x = torch.randn(M, K, device="cuda")
    g = torch.ones(K, device="cuda"); b = torch.zeros(K, device="cuda")
    w32 = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    for nseg in (1, 3, 2):
        np_ = 1 if nseg == 1 else 2
        fmt = {1: 0, 3: 1, 2: 2}[nseg]
        a = torch.empty(M, K * np_, device="cuda", dtype=torch.float16)
        _lib.check(lib.pg_layernorm_f16(x.data_ptr(), K, g.data_ptr(), b.data_ptr(), M, K, a.data_ptr(), K * np_, K if np_ == 2 else 0,
                                        fmt if fmt else 0, 4.0, None))
        w = torch.empty(N, K * np_, device="cuda", dtype=torch.float16)
        winv = torch.ones(N, device="cuda")
        _lib.check(lib.pg_pack_weight(w32.data_ptr(), N, K, fmt, w.data_ptr(), winv.data_ptr(), None))
        args = _lib.PgGemmArgs()
        args.a, args.lda, args.w, args.ldw, args.bias = a.data_ptr(), K * np_, w.data_ptr(), K * np_, bias.data_ptr()
        args.M, args.N, args.K, args.nseg, args.epi = M, N, K, nseg, epi
        args.a_scale, args.w_inv = 4.0, winv.data_ptr()
        if epi == 2:
            res = torch.zeros(M, N, device="cuda"); args.resid, args.ldr = res.data_ptr(), N
        else:
            out = torch.empty(M, N * np_, device="cuda", dtype=torch.float16)
            args.out_h, args.ldo, args.out_lo_off = out.data_ptr(), N * np_, (N if np_ == 2 else 0)
            # QKV feeds the attention kernel (fp16 hi/lo), fc1 feeds fc2 (operand format of the mode)
            args.out_fmt, args.out_scale = ((2 if nseg == 2 else 0) if name == "fc1" else 0), 2.0
        for kc, cta2 in [(k, c) for k in KCH for c in (0, 1)]:
            lib.pg_set_tuning(b"gemm_kchunk", kc)
            lib.pg_set_tuning(b"gemm_cta2", cta2)
            for _ in range(3):
                _lib.check(lib.pg_gemm(C.byref(args), None))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                lib.pg_gemm(C.byref(args), None)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            alg = 2.0 * M * N * K / ms / 1e9
            print(json.dumps({"gemm": name, "N": N, "K": K, "epi": epi, "nseg": nseg, "kchunk": kc, "cta2": cta2, "ms": round(ms, 4),
                              "algorithmic_tflops": round(alg, 1), "issued_tflops": round(alg * {1: 1, 3: 3, 2: 2}[nseg], 1),
                              "issued_frac_of_burst_peak": round(alg * {1: 1, 3: 3, 2: 2}[nseg] / PEAK, 3)}), flush=True)
        lib.pg_set_tuning(b"gemm_cta2", 0)
        del a, w
    del x, w32
