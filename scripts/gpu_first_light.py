"""First-light diagnostics on a B200: run each kernel against torch and print error statistics (no asserts).
Usage: python scripts/gpu_first_light.py [stage]   — without a stage, runs every stage in a subprocess with a timeout."""
import ctypes as C, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

STAGES = ["gemm_small", "gemm_epi", "gemm_x3", "gemm_big", "ln", "attn", "attn_x3", "attn_perf", "model_esm1v", "model_esm2", "model_x3"]
if os.environ.get("PG_STAGES"): STAGES = os.environ["PG_STAGES"].split(",")

def hilo(t):
    hi = t.to(torch.float16); lo = (t - hi.float()).to(torch.float16); return torch.cat([hi, lo], dim=1).contiguous()

def run_gemm(M, N, K, nseg, epi, seed=0, rot=False):
    from proteingym_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    if nseg == 3: a16, w16 = hilo(A), hilo(W)
    else: a16, w16 = A.half().contiguous(), W.half().contiguous()
    Aeff = (a16[:, :K].double() + (a16[:, K:].double() if nseg == 3 else 0)); Weff = (w16[:, :K].double() + (w16[:, K:].double() if nseg == 3 else 0))
    ref = Aeff @ Weff.T + bias.double()
    if nseg == 3:  # the kernel drops lo*lo
        ref = ref - a16[:, K:].double() @ w16[:, K:].double().T
    args = _lib.PgGemmArgs(); args.a = a16.data_ptr(); args.lda = a16.shape[1]; args.w = w16.data_ptr(); args.ldw = w16.shape[1]
    args.bias = bias.data_ptr(); args.M, args.N, args.K, args.nseg, args.epi = M, N, K, nseg, epi
    np_ = 2 if nseg == 3 else 1
    if epi == 2:
        resid = torch.randn(M, N, device="cuda", generator=g); r0 = resid.clone()
        args.resid = resid.data_ptr(); args.ldr = N
        ref = ref + r0.double()
    else:
        out = torch.zeros(M, N * np_, device="cuda", dtype=torch.float16)
        args.out_h = out.data_ptr(); args.ldo = N * np_; args.out_lo_off = N if nseg == 3 else 0
        if epi == 1: ref = ref * 0.5 * (1 + torch.erf(ref / 2 ** 0.5))
    if epi == 3:
        T = 37; cos = torch.rand(T, 32, device="cuda", generator=g); sin = torch.rand(T, 32, device="cuda", generator=g)
        args.rot_cos, args.rot_sin, args.rot_T, args.rot_dim = cos.data_ptr(), sin.data_ptr(), T, (N // 3 // 64) * 64
        d = args.rot_dim; t = torch.arange(M, device="cuda") % T
        c = cos[t].double(); s = sin[t].double()
        r = ref.clone()
        for h0 in range(0, 2 * d, 64):
            x1, x2 = ref[:, h0:h0 + 32], ref[:, h0 + 32:h0 + 64]
            r[:, h0:h0 + 32] = x1 * c - x2 * s; r[:, h0 + 32:h0 + 64] = x2 * c + x1 * s
        ref = r
    torch.cuda.synchronize(); t0 = time.time()
    rc = lib.pg_gemm(C.byref(args), None); torch.cuda.synchronize()
    if rc: print("  rc", rc, lib.pg_last_error(None)); return
    got = resid.double() if epi == 2 else (out[:, :N].double() + (out[:, N:].double() if nseg == 3 else 0))
    err = (got - ref).abs()
    print(f"  gemm M={M} N={N} K={K} nseg={nseg} epi={epi}: max|err|={err.max().item():.3e} mean={err.mean().item():.3e} ref_absmax={ref.abs().max().item():.2f} ({(time.time()-t0)*1e3:.1f} ms incl launch)")
    if err.max().item() > 0.05:
        bad = (err > 0.05).nonzero()
        print("   first bad idx:", bad[:8].tolist(), " n_bad", len(bad), "of", err.numel())
        rows_bad = torch.unique(bad[:, 0]); cols_bad = torch.unique(bad[:, 1])
        print("   bad rows range", rows_bad.min().item(), rows_bad.max().item(), "n", len(rows_bad), " bad cols range", cols_bad.min().item(), cols_bad.max().item(), "n", len(cols_bad))
        print("   got[0,:8]", got[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())

def stage_gemm_small():
    run_gemm(128, 256, 64, 1, 0); run_gemm(128, 256, 256, 1, 0); run_gemm(256, 512, 1280, 1, 0)
def stage_gemm_epi():
    run_gemm(200, 192, 128, 1, 0); run_gemm(300, 320, 128, 1, 1); run_gemm(129, 100, 64, 1, 2); run_gemm(500, 384, 128, 1, 3)
def stage_gemm_x3():
    run_gemm(256, 512, 256, 3, 0); run_gemm(300, 320, 128, 3, 1); run_gemm(300, 320, 128, 3, 2); run_gemm(300, 384, 128, 3, 3)
def stage_gemm_big():
    from proteingym_b200 import _lib
    lib = _lib.load()
    for (M, N, K, nseg, epi) in ((65536, 3840, 1280, 1, 0), (65536, 5120, 1280, 1, 1), (65536, 1280, 5120, 1, 2), (65536, 1280, 1280, 1, 2), (65536, 3840, 1280, 3, 0), (65536, 5120, 1280, 3, 1), (65536, 1280, 5120, 3, 2)):
        np_ = 2 if nseg == 3 else 1
        a = torch.randn(M, K * np_, device="cuda").half(); w = torch.randn(N, K * np_, device="cuda").half(); b = torch.zeros(N, device="cuda")
        out = torch.empty(M, N * np_, device="cuda", dtype=torch.float16)
        args = _lib.PgGemmArgs(); args.a = a.data_ptr(); args.lda = K * np_; args.w = w.data_ptr(); args.ldw = K * np_; args.bias = b.data_ptr()
        args.M, args.N, args.K, args.nseg, args.epi = M, N, K, nseg, epi; args.out_h = out.data_ptr(); args.ldo = N * np_; args.out_lo_off = N if nseg == 3 else 0
        if epi == 2:
            res = torch.zeros(M, N, device="cuda"); args.resid = res.data_ptr(); args.ldr = N
        for _ in range(3): lib.pg_gemm(C.byref(args), None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): lib.pg_gemm(C.byref(args), None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  gemm {M}x{N}x{K} nseg={nseg} epi={epi}: {ms:.3f} ms  {2*M*N*K*nseg/ms/1e9:.1f} TFLOP/s issued")
        a16 = a[:, :K]; w16 = w[:, :K]
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3): torch.matmul(a16, w16.T)
        t0.record()
        for _ in range(10): torch.matmul(a16, w16.T)
        t1.record(); torch.cuda.synchronize()
        print(f"     cuBLAS fp16 same shape: {t0.elapsed_time(t1)/10:.3f} ms  {2*M*N*K/(t0.elapsed_time(t1)/10)/1e9:.1f} TFLOP/s")

def stage_ln():
    from proteingym_b200 import _lib
    lib = _lib.load()
    for rows, d in ((100, 64), (1000, 1280), (77, 2560)):
        x = torch.randn(rows, d, device="cuda") * 3 + 1; g = torch.randn(d, device="cuda"); b = torch.randn(d, device="cuda")
        out = torch.zeros(rows, 2 * d, device="cuda", dtype=torch.float16)
        rc = lib.pg_layernorm_f16(x.data_ptr(), d, g.data_ptr(), b.data_ptr(), rows, d, out.data_ptr(), 2 * d, d, None); torch.cuda.synchronize()
        ref = torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double(), 1e-5)
        got = out[:, :d].double() + out[:, d:].double()
        print(f"  ln rows={rows} d={d} rc={rc}: max|err|={(got-ref).abs().max().item():.3e}  hi-only err={(out[:, :d].double()-ref).abs().max().item():.3e}")

def run_attn(B, T, H, nseg, causal=0, impl=0):
    from proteingym_b200 import _lib
    lib = _lib.load()
    d = H * 64; g = torch.Generator(device="cuda").manual_seed(1)
    qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g); qkv[:, :d] *= 0.3
    q16 = hilo(qkv) if nseg == 3 else qkv.half().contiguous()
    np_ = 2 if nseg == 3 else 1
    eff = q16[:, :3 * d].double() + (q16[:, 3 * d:].double() if nseg == 3 else 0)
    q, k, v = [eff[:, i * d:(i + 1) * d].view(B, T, H, 64).transpose(1, 2) for i in range(3)]
    s = q @ k.transpose(-1, -2)
    if causal: s = s.masked_fill(torch.triu(torch.ones(T, T, device="cuda", dtype=torch.bool), 1), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, d)
    out = torch.zeros(B * T, d * np_, device="cuda", dtype=torch.float16)
    a = _lib.PgAttnArgs(); a.qkv = q16.data_ptr(); a.ld = 3 * d * np_; a.lo_off = 3 * d if nseg == 3 else 0
    a.out = out.data_ptr(); a.ldo = d * np_; a.out_lo_off = d if nseg == 3 else 0
    a.B, a.T, a.heads, a.nseg, a.causal, a.impl = B, T, H, nseg, causal, impl
    rc = lib.pg_attention(C.byref(a), None); torch.cuda.synchronize()
    if rc: print("  rc", rc, lib.pg_last_error(None)); return
    got = out[:, :d].double() + (out[:, d:].double() if nseg == 3 else 0)
    err = (got - ref).abs()
    print(f"  attn impl={impl} B={B} T={T} H={H} nseg={nseg} causal={causal} rc={rc}: max|err|={err.max().item():.3e} mean={err.mean().item():.3e} nan={torch.isnan(got).sum().item()}")
    if err.max().item() > 1e-2 or torch.isnan(got).any():
        e2 = err.view(B, T, H, 64)
        print("   err by query-tile:", [round(e2[:, i:i+128].max().item(), 4) for i in range(0, T, 128)], " by head-dim half:", e2[..., :32].max().item(), e2[..., 32:].max().item())
        print("   got[0,:6]", got[0, :6].tolist(), "ref", ref[0, :6].tolist())

def stage_attn():
    run_attn(2, 64, 2, 1, impl=1); run_attn(2, 130, 2, 1, causal=1, impl=1)
    for (B, T, H) in ((1, 16, 1), (2, 64, 2), (1, 128, 1), (3, 100, 2), (2, 200, 1), (2, 514, 4), (1, 1024, 2), (40, 514, 20)): run_attn(B, T, H, 1, impl=2)
def stage_attn_x3():
    for (B, T, H) in ((1, 16, 1), (2, 64, 2), (3, 100, 2), (2, 514, 4), (40, 514, 20)): run_attn(B, T, H, 3, impl=2)
def stage_attn_perf():
    from proteingym_b200 import _lib
    lib = _lib.load()
    B, T, H = 128, 514, 20; d = H * 64
    for nseg in (1, 3):
        np_ = 2 if nseg == 3 else 1
        qkv = (torch.randn(B * T, 3 * d * np_, device="cuda") * 0.5).half(); out = torch.empty(B * T, d * np_, device="cuda", dtype=torch.float16)
        for impl in (1, 2, 3):
            a = _lib.PgAttnArgs(); a.qkv = qkv.data_ptr(); a.ld = 3 * d * np_; a.lo_off = 3 * d if nseg == 3 else 0
            a.out = out.data_ptr(); a.ldo = d * np_; a.out_lo_off = d if nseg == 3 else 0
            a.B, a.T, a.heads, a.nseg, a.causal, a.impl = B, T, H, nseg, 0, impl
            for _ in range(2): lib.pg_attention(C.byref(a), None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(5): lib.pg_attention(C.byref(a), None)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"  attn perf impl={impl} nseg={nseg} B={B} T={T} H={H}: {ms:.3f} ms  {4*B*H*T*T*64/ms/1e9:.1f} algorithmic TFLOP/s")

def run_model(kind, precision, L=70, layers=2, d=128, heads=2, ffn=256, lnb=False):
    from proteingym_b200 import synth, checkpoint
    from proteingym_b200.esm_engine import EsmScorer
    from oracle import esm_oracle as O
    arch = synth.EsmArch(kind, layers, d, heads, ffn, emb_layer_norm_before=lnb)
    st = synth.make_esm_state(arch, seed=3)
    sc = EsmScorer(checkpoint.config_from_synth(arch), checkpoint.normalise_synth_state(arch, st), precision=precision, max_rows=16384)
    seq = synth.random_protein(L, 11)
    table = sc.masked_marginal_table(seq).cpu().double()
    ref = O.masked_marginal_table(O.load_state(st, kind, torch.float64), seq, kind, layers, heads, dtype=torch.float64, positions=range(1, L + 1))
    err = (table[1:L + 1] - ref[1:L + 1]).abs()
    print(f"  model {kind} {precision} L={L} layers={layers} d={d}: max|dlogp|={err.max().item():.3e} mean={err.mean().item():.3e}  nan={torch.isnan(table[1:L+1]).sum().item()}")
    muts = synth.sample_mutants(seq, 200, 5, multi_frac=0.3)
    got = sc.score_mutants(sc.masked_marginal_table(seq), muts, seq).cpu().double().numpy()
    want = O.score_mutants(muts, seq, ref)
    print(f"     scores: max|d|={np.abs(got-want).max():.3e}  std={want.std():.2f}")
    sc.close()

def stage_model_esm1v(): run_model("esm1v", "f16"); run_model("esm1v", "f16", L=130, lnb=True)
def stage_model_esm2(): run_model("esm2", "f16")
def stage_model_x3(): run_model("esm1v", "f16x3"); run_model("esm2", "f16x3", L=100)
def stage_score(): pass

if __name__ == "__main__":
    if len(sys.argv) > 1:
        torch.manual_seed(0)
        globals()["stage_" + sys.argv[1]]()
    else:
        for st in STAGES:
            print(f"== {st}", flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, st], timeout=240, capture_output=True, text=True)
                print(r.stdout[-4000:], end="")
                if r.returncode: print("  [exit", r.returncode, "]", r.stderr[-1500:])
            except subprocess.TimeoutExpired as e:
                print("  [TIMEOUT]", (e.stdout or b"")[-2000:] if isinstance(e.stdout, (bytes, str)) else "")
