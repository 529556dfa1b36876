"""Run the attention parity cases of tests/test_gpu_parity.py (fp64 reference; causal + ALiBi included) and the throughput probe
for ONE implementation id of pg_attention (0 the model kernel: tcgen05 2 CTAs per SM; 1 mma.sync cross-check). Used to qualify an experimental
kernel without putting it in the test suite: run it under `timeout`. Exit code 0 only if every case passes.
    python scripts/check_attention_impl.py 4"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from proteingym_b200 import _lib  # noqa: E402
import test_gpu_parity as TP  # noqa: E402

CASES = [(2, 64, 2, 1, 0), (3, 100, 2, 1, 0), (2, 514, 4, 1, 0), (1, 1024, 2, 1, 0), (1, 1, 1, 1, 0), (2, 3, 1, 3, 0), (3, 100, 2, 3, 0),
         (2, 514, 4, 3, 0), (2, 130, 2, 1, 1), (2, 130, 2, 3, 1), (2, 514, 4, 1, 1), (1, 1024, 2, 3, 1),
         (1, 128, 1, 1, 0), (1, 129, 1, 3, 0), (2, 384, 2, 3, 0), (2, 385, 2, 1, 1), (3, 700, 3, 3, 1), (40, 514, 20, 3, 0), (40, 514, 20, 1, 0)]


def main():
    impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ok = True
    for (B, T, H, nseg, causal) in CASES:
        try:
            TP.test_attention_matches_fp64.__wrapped__(B, T, H, nseg, causal, impl) if hasattr(TP.test_attention_matches_fp64, "__wrapped__") \
                else TP.test_attention_matches_fp64(B, T, H, nseg, causal, impl)
            print(f"  ok   impl={impl} B={B} T={T} H={H} nseg={nseg} causal={causal}", flush=True)
        except AssertionError as e:
            ok = False
            print(f"  FAIL impl={impl} B={B} T={T} H={H} nseg={nseg} causal={causal}: {str(e)[:200]}", flush=True)
    lib = _lib.load()
    B, T, H = 128, 514, 20
    d = H * 64
    for nseg in (1, 3):
        np_ = 2 if nseg == 3 else 1
        x32 = torch.randn(B * T, 3 * d, device="cuda") * 0.5
        hi16 = x32.half()
        qkv = torch.cat([hi16, (x32 - hi16.float()).half()], 1).contiguous() if np_ == 2 else hi16  # lo plane = fp16 remainder, as in the model
        out = torch.empty(B * T, d * np_, device="cuda", dtype=torch.float16)
        for im in (1, impl):
            a = _lib.PgAttnArgs()
            a.qkv, a.ld, a.lo_off = qkv.data_ptr(), 3 * d * np_, (3 * d if nseg == 3 else 0)
            a.out, a.ldo, a.out_lo_off = out.data_ptr(), d * np_, (d if nseg == 3 else 0)
            a.B, a.T, a.heads, a.nseg, a.causal, a.impl = B, T, H, nseg, 0, im
            for _ in range(3):
                _lib.check(lib.pg_attention(C.byref(a), None))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.pg_attention(C.byref(a), None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"  perf impl={im} nseg={nseg} B={B} T={T} H={H}: {ms:.3f} ms  {4 * B * H * T * T * 64 / ms / 1e9:.1f} algorithmic TFLOP/s", flush=True)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
