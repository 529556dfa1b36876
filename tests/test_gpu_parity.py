"""GPU (-m gpu): parity of the CUDA path, called through the C-ABI, against
  (1) the oracle on seeded inputs, (2) golden vectors produced by the unmodified reference, and
  (3) size-independent properties at BASELINE.json's full size.
Tolerances: per-mutant score |diff| <= 1e-3 in parity mode (f16x3) — north_star's bar; the single-pass f16 mode is only
held to Spearman >= 0.999 and a loose absolute bound."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_SMALL, GOLDEN_WINDOW, load_golden
from oracle import esm_oracle as O
from proteingym_b200 import _lib, checkpoint, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3
PARITY_MODES = ["f16f8", "f16x3"]  # both must meet the 1e-3 bar; f16f8 (2 tensor-pipe units) is the product default


def scorer(arch, state, precision="f16f8", max_rows=32768):
    from proteingym_b200.esm_engine import EsmScorer
    return EsmScorer(checkpoint.config_from_synth(arch), checkpoint.normalise_synth_state(arch, state), precision=precision,
                     max_rows=max_rows)


def kind(arch):
    return "esm2" if arch.kind == "esm2" else "esm1v"


def hilo(t):
    hi = t.to(torch.float16)
    return torch.cat([hi, (t - hi.float()).to(torch.float16)], dim=1).contiguous()


def q8(t):
    """round-to-nearest e4m3 with saturation (what cvt.rn.satfinite.e4m3x2.f32 does), as raw bytes."""
    return t.float().clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)


def dq8(b):
    return b.view(torch.float8_e4m3fn).double()


def pack_a_f8(t, s):
    """fp32 [M, K] -> the nseg-2 `a` operand rows [hi fp16 (2K bytes) | lo8 (K) | hi8 (K)] as uint8 [M, 4K] (common.h fmt 2)."""
    hi = t.to(torch.float16)
    lo = t - hi.float()
    return torch.cat([hi.view(torch.uint8), q8(lo * (2048.0 * s)), q8(hi.float() * s)], dim=1).contiguous()


def unpack_f8(buf, n, s):
    """uint8 [M, 4n] fmt-2 rows -> (hi, lo8 / (2048 s), hi8 / s) as float64 [M, n]."""
    hi = buf[:, :2 * n].contiguous().view(torch.float16).double()
    return hi, dq8(buf[:, 2 * n:3 * n].contiguous()) / (2048.0 * s), dq8(buf[:, 3 * n:].contiguous()) / s


def spearman(a, b):
    ra = np.argsort(np.argsort(a)).astype(np.float64)
    rb = np.argsort(np.argsort(b)).astype(np.float64)
    return np.corrcoef(ra, rb)[0, 1]


# ------------------------------------------------------------------------------------------------ kernel level
@pytest.mark.parametrize("M,N,K,nseg,epi", [(128, 256, 64, 1, 0), (1, 8, 64, 1, 0), (257, 520, 192, 1, 1), (129, 100, 64, 1, 2),
                                            (500, 384, 128, 1, 3), (300, 320, 128, 3, 0), (300, 320, 128, 3, 1),
                                            (1000, 1280, 1280, 3, 2), (300, 384, 128, 3, 3), (4096, 3840, 1280, 1, 0), (300, 320, 128, 3, 4)])
@pytest.mark.parametrize("cta2", [0, 1])
def test_gemm_matches_fp64(M, N, K, nseg, epi, cta2):
    lib = _lib.load()
    lib.pg_set_tuning(b"gemm_cta2", cta2)  # 1: the CTA-pair (cta_group::2) form of the kernel
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    a16, w16 = (hilo(A), hilo(W)) if nseg == 3 else (A.half().contiguous(), W.half().contiguous())
    Ah, Wh = a16[:, :K].double(), w16[:, :K].double()
    ref = Ah @ Wh.T + bias.double()
    if nseg == 3:
        ref = ref + a16[:, K:].double() @ Wh.T + Ah @ w16[:, K:].double().T
    args = _lib.PgGemmArgs()
    args.a, args.lda, args.w, args.ldw, args.bias = a16.data_ptr(), a16.shape[1], w16.data_ptr(), w16.shape[1], bias.data_ptr()
    args.M, args.N, args.K, args.nseg, args.epi = M, N, K, nseg, epi
    npl = 2 if nseg == 3 else 1
    if epi == 2:
        resid = torch.randn(M, N, device="cuda", generator=g)
        ref = ref + resid.double()
        args.resid, args.ldr = resid.data_ptr(), N
    else:
        out = torch.zeros(M, N * npl, device="cuda", dtype=torch.float16)
        args.out_h, args.ldo, args.out_lo_off = out.data_ptr(), N * npl, (N if nseg == 3 else 0)
        if epi == 1:
            ref = ref * 0.5 * (1 + torch.erf(ref / 2 ** 0.5))
        if epi == 4:
            ref = torch.relu(ref) ** 2  # tranception/activations.py:79-84
    if epi == 3:
        T = 37
        cos, sin = torch.rand(T, 32, device="cuda", generator=g), torch.rand(T, 32, device="cuda", generator=g)
        d = (N // 3 // 64) * 64
        args.rot_cos, args.rot_sin, args.rot_T, args.rot_dim = cos.data_ptr(), sin.data_ptr(), T, d
        t = torch.arange(M, device="cuda") % T
        c, s = cos[t].double(), sin[t].double()
        r = ref.clone()
        for h0 in range(0, 2 * d, 64):
            x1, x2 = ref[:, h0:h0 + 32], ref[:, h0 + 32:h0 + 64]
            r[:, h0:h0 + 32], r[:, h0 + 32:h0 + 64] = x1 * c - x2 * s, x2 * c + x1 * s
        ref = r
    try:
        _lib.check(lib.pg_gemm(C.byref(args), None))
        torch.cuda.synchronize()
    finally:
        lib.pg_set_tuning(b"gemm_cta2", 1)  # back to the default (CTA pairs) for the model-level tests that follow
    got = resid.double() if epi == 2 else out[:, :N].double() + (out[:, N:].double() if nseg == 3 else 0)
    err = (got - ref).abs().max().item()
    # fp32 accumulate of exactly-representable products; the fp16 output rounding dominates for single-plane outputs
    amax = ref.abs().max().item()
    bound = 4e-7 * (K * nseg) ** 0.5 * max(4.0, amax) if (epi == 2 or nseg == 3) else 2.5e-3 * max(1.0, amax / 4)
    assert err < bound, (err, bound)


@pytest.mark.parametrize("M,N,K,epi,out_fmt", [(128, 256, 64, 0, 1), (300, 320, 128, 1, 2), (257, 576, 192, 2, 0), (1000, 1280, 1280, 2, 0),
                                               (500, 384, 128, 3, 1), (640, 5120, 1280, 1, 2), (300, 1280, 5120, 2, 0),
                                               (130, 192, 10240, 0, 2)])
@pytest.mark.parametrize("cta2", [0, 1])
def test_gemm_f16f8_matches_fp64(M, N, K, epi, out_fmt, cta2):
    """nseg 2: fp16 hi*hi + e4m3 cross terms. The kernel's products are exact in fp32, so against an fp64 evaluation of the SAME
    quantised planes only accumulation error remains; pg_pack_weight's row scales are checked on the way."""
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    W[::7] *= 37.0  # rows of very different magnitude: e4m3 is floating point, one scale per matrix has to be enough
    bias = torch.randn(N, device="cuda", generator=g)
    sA, sO = 4.0, 2.0
    a_op = pack_a_f8(A, sA)
    w_op = torch.zeros(N, 4 * K, device="cuda", dtype=torch.uint8)
    w_inv = torch.zeros(N, device="cuda")
    _lib.check(lib.pg_pack_weight(W.data_ptr(), N, K, 2, w_op.data_ptr(), w_inv.data_ptr(), None))
    torch.cuda.synchronize()
    t = 1.0 / w_inv.double()
    whi = w_op[:, :2 * K].contiguous().view(torch.float16)
    assert torch.equal(whi, W.to(torch.float16))
    amax = whi.float().abs().max().double()  # ONE power-of-two scale per matrix: its largest |hi| lands in (112, 224]
    mant, _ = torch.frexp(t)
    assert torch.all(mant == 0.5) and torch.all(t == t[0]) and amax * t[0] <= 224.0 and amax * t[0] > 112.0
    assert torch.equal(w_op[:, 2 * K:3 * K], q8(whi.float() * t[:, None].float()))
    assert torch.equal(w_op[:, 3 * K:], q8((W - whi.float()) * (2048.0 * t[:, None].float())))
    ahi, alo8, ahi8 = unpack_f8(a_op, K, sA)
    whi8 = dq8(w_op[:, 2 * K:3 * K].contiguous()) / t[:, None]
    wlo8 = dq8(w_op[:, 3 * K:].contiguous()) / (2048.0 * t[:, None])
    ref = ahi @ whi.double().T + alo8 @ whi8.T + ahi8 @ wlo8.T + bias.double()
    full = A.double() @ W.double().T + bias.double()
    args = _lib.PgGemmArgs()
    args.a, args.lda, args.w, args.ldw, args.bias = a_op.data_ptr(), 2 * K, w_op.data_ptr(), 2 * K, bias.data_ptr()
    args.M, args.N, args.K, args.nseg, args.epi = M, N, K, 2, epi
    args.a_scale, args.w_inv = sA, w_inv.data_ptr()
    if epi == 2:
        resid = torch.randn(M, N, device="cuda", generator=g)
        ref = ref + resid.double()
        full = full + resid.double()
        args.resid, args.ldr = resid.data_ptr(), N
    else:
        out = torch.zeros(M, 4 * N, device="cuda", dtype=torch.uint8)
        args.out_h, args.ldo, args.out_lo_off, args.out_fmt, args.out_scale = out.data_ptr(), 2 * N, N, out_fmt, sO
        if epi == 1:
            ref = ref * 0.5 * (1 + torch.erf(ref / 2 ** 0.5))
            full = full * 0.5 * (1 + torch.erf(full / 2 ** 0.5))
    if epi == 3:
        T = 37
        cos, sin = torch.rand(T, 32, device="cuda", generator=g), torch.rand(T, 32, device="cuda", generator=g)
        d = (N // 3 // 64) * 64
        args.rot_cos, args.rot_sin, args.rot_T, args.rot_dim = cos.data_ptr(), sin.data_ptr(), T, d
        tt = torch.arange(M, device="cuda") % T
        c, sn = cos[tt].double(), sin[tt].double()

        def rot(x):
            r = x.clone()
            for h0 in range(0, 2 * d, 64):
                x1, x2 = x[:, h0:h0 + 32], x[:, h0 + 32:h0 + 64]
                r[:, h0:h0 + 32], r[:, h0 + 32:h0 + 64] = x1 * c - x2 * sn, x2 * c + x1 * sn
            return r
        ref, full = rot(ref), rot(full)
    lib.pg_set_tuning(b"gemm_cta2", cta2)
    try:
        _lib.check(lib.pg_gemm(C.byref(args), None))
        torch.cuda.synchronize()
    finally:
        lib.pg_set_tuning(b"gemm_cta2", 1)
    amax = ref.abs().max().item()
    bound = 4e-7 * (3 * K) ** 0.5 * max(4.0, amax)
    if epi == 2:
        got = resid.double()
    elif out_fmt == 1:
        o16 = out.view(torch.float16)
        got = o16[:, :N].double() + o16[:, N:].double()
    else:
        hi, lo8, hi8 = unpack_f8(out, N, sO)
        got = hi + lo8
        # the e4m3 planes are what the next GEMM's cross terms read: hi8 ~ hi within e4m3 rounding, lo8 ~ (x - hi) likewise
        assert ((hi8 - hi).abs() <= 2 ** -4 * hi.abs() + 2 ** -10 / sO).all()
        assert torch.equal(out[:, 3 * N:], q8(hi.float() * sO))
        bound = bound + 2 ** -16 * max(1.0, amax)  # lo carried at 4 bits instead of 11
    err = (got - ref).abs().max().item()
    print(f"\nf16f8 gemm {M}x{N}x{K} epi {epi}: |got - planes_fp64| = {err:.2e} (bound {bound:.2e}); |got - exact fp32-operand product| = "
          f"{(got - full).abs().max().item():.2e}")
    assert err < bound, (err, bound)


def test_gemm_rejects_bad_arguments():
    lib = _lib.load()
    a = torch.zeros(8, 60, device="cuda", dtype=torch.float16)
    args = _lib.PgGemmArgs()
    args.a, args.lda, args.w, args.ldw, args.M, args.N, args.K, args.nseg, args.epi = a.data_ptr(), 60, a.data_ptr(), 60, 8, 8, 60, 1, 0
    args.out_h, args.ldo = a.data_ptr(), 60
    assert lib.pg_gemm(C.byref(args), None) == 1 and b"multiple of 64" in lib.pg_last_error(None)
    args.K = 64
    args.nseg = 4
    assert lib.pg_gemm(C.byref(args), None) == 1
    args.nseg = 2  # fp16 + e4m3 mode without the weight scales
    assert lib.pg_gemm(C.byref(args), None) == 1 and b"w_inv" in lib.pg_last_error(None)


@pytest.mark.parametrize("rows,d", [(1, 64), (100, 64), (1000, 1280), (77, 2560)])
def test_layernorm_matches_fp64(rows, d):
    lib = _lib.load()
    x = torch.randn(rows, d, device="cuda") * 3 + 1
    g, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    out = torch.zeros(rows, 2 * d, device="cuda", dtype=torch.float16)
    _lib.check(lib.pg_layernorm_f16(x.data_ptr(), d, g.data_ptr(), b.data_ptr(), rows, d, out.data_ptr(), 2 * d, d, 1, 0.0, None))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double(), 1e-5)
    assert (out[:, :d].double() + out[:, d:].double() - ref).abs().max().item() < 1e-5
    # fmt 2: same hi plane, e4m3 planes consistent with it
    out8 = torch.zeros(rows, 4 * d, device="cuda", dtype=torch.uint8)
    _lib.check(lib.pg_layernorm_f16(x.data_ptr(), d, g.data_ptr(), b.data_ptr(), rows, d, out8.data_ptr(), 2 * d, d, 2, 4.0, None))
    torch.cuda.synchronize()
    hi, lo8, hi8 = unpack_f8(out8, d, 4.0)
    assert torch.equal(out8[:, :2 * d].contiguous().view(torch.float16), out[:, :d])
    assert torch.equal(out8[:, 3 * d:], q8(out[:, :d].float() * 4.0))
    assert (hi + lo8 - ref).abs().max().item() < 2 ** -15 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,T,H,nseg,causal", [(2, 64, 2, 1, 0), (3, 100, 2, 1, 0), (2, 514, 4, 1, 0), (1, 1024, 2, 1, 0),
                                               (1, 1, 1, 1, 0), (2, 3, 1, 3, 0), (3, 100, 2, 3, 0), (2, 514, 4, 3, 0),
                                               (2, 130, 2, 1, 1), (2, 130, 2, 3, 1), (2, 514, 4, 1, 1), (1, 1024, 2, 3, 1)])
@pytest.mark.parametrize("impl", [0, 1])
def test_attention_matches_fp64(B, T, H, nseg, causal, impl):
    slopes = None
    lib = _lib.load()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * T + H)
    qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g)
    qkv[:, :d] *= 0.3
    q16 = hilo(qkv) if nseg == 3 else qkv.half().contiguous()
    npl = 2 if nseg == 3 else 1
    eff = q16[:, :3 * d].double() + (q16[:, 3 * d:].double() if nseg == 3 else 0)
    q, k, v = [eff[:, i * d:(i + 1) * d].view(B, T, H, 64).transpose(1, 2) for i in range(3)]
    s = q @ k.transpose(-1, -2)
    if causal:
        slopes = torch.tensor([2.0 ** -(i + 3) for i in range(H)], device="cuda", dtype=torch.float32)  # ALiBi rides with causal
        s = s + slopes.double()[None, :, None, None] * torch.arange(T, device="cuda", dtype=torch.float64)[None, None, None, :]
        s = s.masked_fill(torch.triu(torch.ones(T, T, device="cuda", dtype=torch.bool), 1), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, d)
    out = torch.zeros(B * T, d * npl, device="cuda", dtype=torch.float16)
    a = _lib.PgAttnArgs()
    a.alibi_slopes = slopes.data_ptr() if slopes is not None else None
    a.qkv, a.ld, a.lo_off = q16.data_ptr(), 3 * d * npl, (3 * d if nseg == 3 else 0)
    a.out, a.ldo, a.out_lo_off = out.data_ptr(), d * npl, (d if nseg == 3 else 0)
    a.B, a.T, a.heads, a.nseg, a.causal, a.impl = B, T, H, nseg, causal, impl
    _lib.check(lib.pg_attention(C.byref(a), None))
    torch.cuda.synchronize()
    got = out[:, :d].double() + (out[:, d:].double() if nseg == 3 else 0)
    assert (got - ref).abs().max().item() < ((1e-4 if causal else 3e-5) if nseg == 3 else 3e-3)
    if impl == 0 and nseg == 3:  # same launch writing the out_proj operand with e4m3 planes (common.h fmt 2)
        out8 = torch.zeros(B * T, 4 * d, device="cuda", dtype=torch.uint8)
        a.out, a.out_fmt, a.out_scale = out8.data_ptr(), 2, 4.0
        _lib.check(lib.pg_attention(C.byref(a), None))
        torch.cuda.synchronize()
        hi, lo8, hi8 = unpack_f8(out8, d, 4.0)
        assert torch.equal(out8[:, :2 * d].contiguous().view(torch.float16), out[:, :d])
        assert torch.equal(out8[:, 3 * d:], q8(out[:, :d].float() * 4.0))
        assert (hi + lo8 - got).abs().max().item() < 2 ** -15 * max(1.0, got.abs().max().item())


def test_score_mutants_bit_exact_vs_label_row():
    lib = _lib.load()
    seq = synth.random_protein(80, 4)
    table = torch.randn(82, 33, device="cuda")
    muts = synth.sample_mutants(seq, 500, 6, multi_frac=0.4)
    from proteingym_b200.mutants import parse_mutants
    r, w, m, o = (torch.from_numpy(x).cuda() for x in parse_mutants(muts, seq))
    out = torch.empty(len(muts), device="cuda")
    _lib.check(lib.pg_score_mutants(table.data_ptr(), 82, 33, r.data_ptr(), w.data_ptr(), m.data_ptr(), o.data_ptr(), len(muts),
                                    out.data_ptr(), None))
    tc = table.cpu()
    want = []
    for mu in muts:  # label_row with fp32 accumulation in site order (what the kernel does)
        s = np.float32(0)
        for site in mu.split(":"):
            i = int(site[1:-1])
            s = np.float32(s + np.float32(tc[i, O.TOK[site[-1]]] - tc[i, O.TOK[site[0]]]))
        want.append(s)
    assert np.array_equal(out.cpu().numpy(), np.asarray(want, dtype=np.float32))
    assert lib.pg_score_mutants(table.data_ptr(), 82, 33, None, None, None, None, 0, None, None) == 0  # empty frame


# ------------------------------------------------------------------------------------------------ model level vs oracle
@pytest.mark.parametrize("knd,L,layers,d,heads,ffn,lnb", [("esm1v", 70, 2, 128, 2, 256, False), ("esm1v", 130, 2, 128, 2, 256, True),
                                                          ("esm2", 100, 3, 128, 2, 512, False), ("esm1v", 37, 1, 64, 1, 64, False),
                                                          ("esm1v", 257, 4, 256, 4, 1024, False)])
@pytest.mark.parametrize("mode", PARITY_MODES)
def test_model_matches_oracle_parity_mode(knd, L, layers, d, heads, ffn, lnb, mode):
    arch = synth.EsmArch(knd, layers, d, heads, ffn, emb_layer_norm_before=lnb)
    st = synth.make_esm_state(arch, seed=3)
    seq = synth.random_protein(L, 11)
    sc = scorer(arch, st, precision=mode)
    table = sc.masked_marginal_table(seq).cpu().double()
    ref = O.masked_marginal_table(O.load_state(st, knd, torch.float64), seq, knd, layers, heads, dtype=torch.float64,
                                  positions=range(1, L + 1))
    assert torch.isnan(table[0]).all() and torch.isnan(table[L + 1]).all()  # BOS/EOS rows are never read by label_row
    assert (table[1:L + 1] - ref[1:L + 1]).abs().max().item() < 2e-4
    muts = synth.sample_mutants(seq, 300, 5, multi_frac=0.3)
    got = sc.score_assay(seq, muts).astype(np.float64)
    assert np.abs(got - O.score_mutants(muts, seq, ref)).max() < TOL
    sc.close()


def test_fast_mode_keeps_rank_order():
    arch = synth.EsmArch("esm1v", 4, 256, 4, 1024)
    st = synth.make_esm_state(arch, seed=3)
    seq = synth.random_protein(120, 11)
    muts = synth.all_single_mutants(seq)
    sc = scorer(arch, st, precision="f16")
    got = sc.score_assay(seq, muts).astype(np.float64)
    sc.close()
    ref = O.masked_marginal_table(O.load_state(st, "esm1v", torch.float64), seq, "esm1v", 4, 4, dtype=torch.float64,
                                  positions=range(1, 121))
    want = O.score_mutants(muts, seq, ref)
    assert spearman(got, want) >= 0.999 and np.abs(got - want).max() < 5e-2


def test_wt_marginals_full_table_matches_oracle():
    arch = synth.EsmArch("esm2", 2, 128, 2, 512)
    st = synth.make_esm_state(arch, seed=5)
    seq = synth.random_protein(90, 1)
    sc = scorer(arch, st)
    got = sc.wt_marginal_table(seq).cpu().double()
    sc.close()
    toks = O.tokenize(seq)[None]
    ref = torch.log_softmax(O.esm_forward(O.load_state(st, "esm2", torch.float64), toks, "esm2", 2, 2, True, torch.float64), -1)[0]
    assert (got - ref).abs().max().item() < 2e-4


def test_error_paths():
    from proteingym_b200.esm_engine import EsmScorer
    arch = synth.EsmArch("esm1v", 1, 64, 1, 64)
    st = checkpoint.normalise_synth_state(arch, synth.make_esm_state(arch))
    bad = dict(st)
    bad.pop("layers.0.fc1.bias")
    with pytest.raises(_lib.PgError, match="fc1.bias"):
        EsmScorer(checkpoint.config_from_synth(arch), bad)
    cfg = checkpoint.config_from_synth(arch)
    cfg.heads = 2  # head_dim 32: unsupported
    with pytest.raises(_lib.PgError, match="head_dim"):
        EsmScorer(cfg, st)
    sc = scorer(arch, synth.make_esm_state(arch))
    with pytest.raises(AssertionError, match="The listed wildtype does not match the provided sequence"):
        sc.score_assay("MKVLAAGIC", ["A1G"])
    assert sc.score_assay("MKVLAAGIC", []).shape == (0,)  # empty frame
    sc.close()


# ------------------------------------------------------------------------------------------------ golden vectors (reference)
@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_golden_small_reference_outputs(name):
    g = load_golden(name)
    arch, seq, df = g["arch"], g["seq"], g["df"]
    sc = scorer(arch, g["state"]())
    L = len(seq)
    table = sc.masked_marginal_table(seq).cpu().numpy()
    assert np.abs(table[1:L + 1] - g["table"][1:L + 1]).max() < 2e-4
    col = g["meta"]["ckpt_names"][0].split(".")[0]
    got = sc.score_assay(seq, list(df["mutant"]))
    assert np.abs(got - df[col].to_numpy()).max() < TOL
    sc.close()


@pytest.mark.parametrize("name", GOLDEN_WINDOW)
def test_golden_windowed_reference_outputs(name):
    # L + 2 = 1102 tokens: every masked row has its own 1024-token window (compute_fitness.py:492-495)
    g = load_golden(name)
    arch, seq, df = g["arch"], g["seq"], g["df"]
    sc = scorer(arch, g["state"](), max_rows=65536)
    col = g["meta"]["ckpt_names"][0].split(".")[0]
    got = sc.score_assay(seq, list(df["mutant"]))
    assert np.abs(got - df[col].to_numpy()).max() < TOL
    pos = [1, 300, 511, 512, 513, 589, 590, 591, 1100]
    rows = sc.masked_marginal_table(seq, positions=pos).cpu().numpy()
    assert np.abs(rows[pos] - g["table"][pos]).max() < 2e-4
    sc.close()


def test_golden_true_size_blat_config1():
    """BASELINE.json config 1: BLAT_ECOLX_Stiffler_2015 (L=286), all 19 substitutions at 24..286, ESM-1v 650M architecture;
    golden = the unmodified reference CLI's CSV (fp32 CPU). Both precision modes."""
    g = load_golden("blat_esm1v_650m")
    arch, seq, df = g["arch"], g["seq"], g["df"]
    col = g["meta"]["ckpt_names"][0].split(".")[0]
    want = df[col].to_numpy()
    st = g["state"]()
    for mode in PARITY_MODES:
        sc = scorer(arch, st, precision=mode, max_rows=131072)
        got = sc.score_assay(seq, list(df["mutant"]))
        err = np.abs(got - want)
        tab = sc.masked_marginal_table(seq, positions=range(24, 287)).cpu().numpy()
        terr = np.abs(tab[24:287] - g["table"][24:287]).max()
        print(f"\nBLAT 650M {mode}: max|d|={err.max():.2e} mean={err.mean():.2e} spearman={spearman(got, want):.6f} max|dlogp|={terr:.2e}")
        assert err.max() < TOL and spearman(got, want) >= 0.999
        assert terr < 5e-4
        sc.close()
    assert np.allclose(df["Ensemble_ESM1v"].to_numpy(), want, atol=1e-6)  # single checkpoint: ensemble == column
    sc = scorer(arch, st, precision="f16", max_rows=131072)
    got = sc.score_assay(seq, list(df["mutant"]))
    err = np.abs(got - want)
    print(f"BLAT 650M f16  : max|d|={err.max():.2e} mean={err.mean():.2e} spearman={spearman(got, want):.6f}")
    assert spearman(got, want) >= 0.999 and err.max() < 0.1
    sc.close()


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_config2():
    """ESM-1v 650M, L=512, 5000 mutants (BASELINE config 2). No CPU oracle at this size; check properties instead:
    determinism, independence from workspace chunking, multi-mutant additivity, log-prob normalisation."""
    arch = synth.ESM1V_650M
    st = synth.make_esm_state(arch, seed=0)
    seq = synth.random_protein(512, 0)
    muts = synth.sample_mutants(seq, 5000, 1000)
    sc = scorer(arch, st, max_rows=131072)
    a = sc.score_assay(seq, muts)
    b = sc.score_assay(seq, muts)
    assert np.array_equal(a, b)  # bit-identical re-run
    table = sc.masked_marginal_table(seq)
    lse = torch.logsumexp(table[1:513], dim=-1)
    assert lse.abs().max().item() < 1e-4  # every emitted row is a normalised log-distribution
    singles = muts[:50]
    pair = [f"{x}:{y}" for x, y in zip(singles[::2], singles[1::2]) if x[1:-1] != y[1:-1]]
    s1 = dict(zip(singles, sc.score_assay(seq, singles)))
    sp = sc.score_assay(seq, pair)
    for pm, v in zip(pair, sp):  # additive over independently masked sites (compute_fitness.py:240-250)
        x, y = pm.split(":")
        assert abs(v - (np.float32(s1[x]) + np.float32(s1[y]))) < 1e-5
    sc.close()
    sc2 = scorer(arch, st, max_rows=40000)  # forces 7 passes of 77 rows instead of 3 of 255
    c = sc2.score_assay(seq, muts)
    sc2.close()
    assert np.array_equal(a, c)  # per-row results do not depend on how rows are batched


# ------------------------------------------------------------------------------------------------ CLI drop-in (seam B1)
@pytest.mark.parametrize("name,model_type", [("tiny_esm1v", "ESM1v"), ("tiny_esm2", "ESM2"), ("tiny_esm1b", "ESM1b")])
def test_cli_output_csv_matches_reference_cli(tmp_path, name, model_type):
    """Run our compute_fitness-compatible CLI on the inputs the reference CLI was run on (oracle/gen_golden.py) and diff
    the CSVs: same columns in the same order, non-score columns identical, score columns within 1e-3."""
    import pandas as pd
    from proteingym_b200 import compute_fitness as cf
    g = load_golden(name)
    arch, seq, ref = g["arch"], g["seq"], g["df"]
    ck = []
    seeds = [g["meta"]["seed"]] + ([g["meta"]["extra_seed"]] if g["meta"]["extra_seed"] is not None else [])
    for fn, seed in zip(g["meta"]["ckpt_names"], seeds):
        synth.write_esm_checkpoint(str(tmp_path / fn), arch, seed=seed)
        ck.append(str(tmp_path / fn))
    (tmp_path / "dms").mkdir()
    ref[["mutant", "mutated_sequence", "DMS_score", "DMS_score_bin"]].to_csv(tmp_path / "dms" / f"{name}.csv", index=False)
    synth.write_mapping_csv(str(tmp_path / "map.csv"), [("OTHER_ASSAY", "other.csv", "MKV"), (name, f"{name}.csv", seq)])
    args = cf.create_parser().parse_args(["--model-location", *ck, "--model_type", model_type, "--dms_index", "1", "--dms_mapping",
                                          str(tmp_path / "map.csv"), "--dms-input", str(tmp_path / "dms"), "--dms-output",
                                          str(tmp_path / "out"), "--scoring-strategy", "masked-marginals", "--scoring-window", "optimal"])
    cf.main(args)
    got = pd.read_csv(tmp_path / "out" / f"{name}.csv")
    assert list(got.columns) == list(ref.columns)
    score_cols = [c.split(".")[0] for c in g["meta"]["ckpt_names"]] + (["Ensemble_ESM1v"] if model_type == "ESM1v" else [])
    for c in ref.columns:
        if c in score_cols:
            assert np.abs(got[c].to_numpy() - ref[c].to_numpy()).max() < TOL, c
        elif ref[c].dtype.kind == "f":
            assert np.allclose(got[c].to_numpy(), ref[c].to_numpy(), rtol=0, atol=1e-12), c
        else:
            assert got[c].equals(ref[c]), c


def test_cli_wt_marginals_and_pseudo_ppl_match_oracle(tmp_path):
    import pandas as pd
    from proteingym_b200 import compute_fitness as cf
    arch = synth.EsmArch("esm2", 2, 128, 2, 512)
    st = synth.write_esm_checkpoint(str(tmp_path / "esm2_t2_x.pt"), arch, seed=4)
    seq = synth.random_protein(40, 8)
    muts = synth.sample_mutants(seq, 30, 2, multi_frac=0.3)
    synth.write_dms_csv(str(tmp_path / "assay.csv"), seq, muts)
    ost = O.load_state(st, "esm2", torch.float64)
    for strategy in ("wt-marginals", "pseudo-ppl"):
        args = cf.create_parser().parse_args(["--model-location", str(tmp_path / "esm2_t2_x.pt"), "--model_type", "ESM2", "--dms-input",
                                              str(tmp_path / "assay.csv"), "--dms-output", str(tmp_path / ("out_" + strategy)),
                                              "--target_seq", seq, "--scoring-strategy", strategy])
        cf.main(args)
        got = pd.read_csv(tmp_path / ("out_" + strategy) / "assay.csv")["esm2_t2_x"].to_numpy()
        if strategy == "wt-marginals":  # compute_fitness.py:475-485
            lp = torch.log_softmax(O.esm_forward(ost, O.tokenize(seq)[None], "esm2", 2, 2, True, torch.float64), -1)[0]
            want = O.score_mutants(muts, seq, lp)
        else:  # compute_pppl, compute_fitness.py:258-279 (incl. its sequence[i]-at-token-i indexing)
            want = []
            for m in muts:
                ms = synth.apply_mutant(seq, m)
                t = O.masked_marginal_table(ost, ms, "esm2", 2, 2, dtype=torch.float64, positions=range(1, len(ms) - 1))
                want.append(sum(t[i, O.TOK[ms[i]]].item() for i in range(1, len(ms) - 1)))
            want = np.asarray(want)
        assert np.abs(got - want).max() < TOL, strategy


def test_wt_marginals_overlapping_windows_match_oracle():
    # compute_fitness.py:435-473 on a 1300-token sequence (two end windows + overlap rule)
    arch = synth.EsmArch("esm1v", 1, 64, 1, 64)
    st = synth.make_esm_state(arch, seed=6)
    seq = synth.random_protein(1298, 3)
    sc = scorer(arch, st)
    got = sc.wt_marginal_table_overlapping(seq).cpu().double()
    sc.close()
    import math
    ost = O.load_state(st, "esm1v", torch.float64)
    toks = O.tokenize(seq)[None]
    n = toks.shape[1]
    probs, wsum = torch.zeros(n, 33, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    w = torch.ones(1024, dtype=torch.float64)
    for i in range(1, 257):
        w[i] = 1 / (1 + math.exp(-(i - 128) / 16))
    for i in range(1022 - 256, 1023):
        w[i] = 1 / (1 + math.exp((i - 1022 + 128) / 16))
    windows = []
    sl, el, sr, er = 0, 1023, (n - 1) - 1024 + 1, n - 1
    while True:
        windows += [sl, sr]
        if el > sr:
            break
        sl += 511; el += 511; sr -= 511; er -= 511
    if el - sr + 1 < 511:
        windows.append(int(n / 2) - 512)
    for s0 in windows:
        lp = torch.log_softmax(O.esm_forward(ost, toks[:, s0:s0 + 1024], "esm1v", 1, 1, True, torch.float64), -1)[0]
        probs[s0:s0 + 1024] += lp * w[:, None]
        wsum[s0:s0 + 1024] += w
    assert (got - probs / wsum[:, None]).abs().max().item() < 2e-4


def test_run_assays_driver_single_gpu_matches_cli_outputs(tmp_path):
    """The multi-assay driver (one process per GPU; here world = 1) writes the same CSVs as per-assay compute_fitness runs."""
    import pandas as pd
    from proteingym_b200 import run_assays
    arch = synth.EsmArch("esm1v", 2, 128, 2, 256)
    st = synth.write_esm_checkpoint(str(tmp_path / "esm1v_tiny_a.pt"), arch, seed=1)
    (tmp_path / "dms").mkdir()
    rows = []
    for k, L in enumerate((40, 90, 25)):
        seq = synth.random_protein(L, 50 + k)
        synth.write_dms_csv(str(tmp_path / "dms" / f"a{k}.csv"), seq, synth.sample_mutants(seq, 30, k, multi_frac=0.2))
        rows.append((f"ASSAY{k}", f"a{k}.csv", seq))
    synth.write_mapping_csv(str(tmp_path / "map.csv"), rows)
    run_assays.main(["--model-location", str(tmp_path / "esm1v_tiny_a.pt"), "--model_type", "ESM1v", "--dms_mapping", str(tmp_path / "map.csv"),
                     "--dms-input", str(tmp_path / "dms"), "--dms-output", str(tmp_path / "out")])
    ost = O.load_state(st, "esm1v", torch.float64)
    for k, (aid, fn, seq) in enumerate(rows):
        got = pd.read_csv(tmp_path / "out" / f"{aid}.csv")
        table = O.masked_marginal_table(ost, seq, "esm1v", 2, 2, dtype=torch.float64)
        want = O.score_mutants(got["mutant"], seq, table)
        assert np.abs(got["esm1v_tiny_a"].to_numpy() - want).max() < TOL
        assert np.allclose(got["Ensemble_ESM1v"], got["esm1v_tiny_a"])
    assert (tmp_path / "out" / "_run_assays_summary.csv").exists()


def test_esm2_3b_true_size_rows_match_oracle():
    """BASELINE config 3 architecture (ESM2 3B: 36 x 2560, 40 heads, ffn 10240, rotary) at true size on a short protein:
    a few masked rows against the fp32 CPU oracle."""
    arch = synth.ESM2_3B
    st = synth.make_esm_state(arch, seed=2)
    seq = synth.random_protein(48, 9)
    pos = [1, 17, 48]
    sc = scorer(arch, st, max_rows=4096)
    got = sc.masked_marginal_table(seq, positions=pos).cpu().numpy()
    sc.close()
    ref = O.masked_marginal_table(O.load_state(st, "esm2"), seq, "esm2", arch.layers, arch.heads, positions=pos, batch=3)
    err = np.abs(got[pos] - ref[pos].numpy()).max()
    aa = [O.TOK[c] for c in synth.AA20]
    ds = lambda t: np.stack([t[i][aa] - t[i][O.TOK[seq[i - 1]]] for i in pos])  # label_row's differences: what a score is made of
    serr = np.abs(ds(got) - ds(ref.numpy())).max()
    print(f"\\nESM2-3B f16x3: max|dlogp| = {err:.2e}, max|d(score term)| = {serr:.2e}")
    assert serr < TOL and err < 3e-3


@pytest.mark.parametrize("mode", ["auto", "f16f8"])
def test_golden_true_size_esm2_3b_multi_mutants(mode):
    """BASELINE config 3 architecture at TRUE SIZE (ESM2 3B: 36 x 2560, 40 heads, ffn 10240, rotary) against the UNMODIFIED reference
    CLI (oracle/gen_golden.py esm2_3b): a 256-residue protein, 300 mutants of which 215 have 2-5 sites (the errors of independently
    masked sites add up), and the reference's own log-prob rows at 42 positions.
    `auto` (what the CLIs run: esm_engine.choose_precision -> f16x3 for a model this wide) must meet the 1e-3 bar on every mutant.
    f16f8 is measured for the record: it meets the bar on single-site mutants and misses it on some 3-5-site ones at this size, which
    is exactly why `auto` does not select it here."""
    from proteingym_b200.esm_engine import choose_precision
    g = load_golden("esm2_3b_multi")
    arch, seq, df = g["arch"], g["seq"], g["df"]
    want = df[g["meta"]["ckpt_names"][0].split(".")[0]].to_numpy()
    prec = choose_precision(checkpoint.config_from_synth(arch), list(df["mutant"])) if mode == "auto" else mode
    assert mode != "auto" or prec == "f16x3"
    sc = scorer(arch, g["state"](), precision=prec, max_rows=32768)
    got = sc.score_assay(seq, list(df["mutant"]))
    pos = g["meta"]["table_positions"]
    tab = sc.masked_marginal_table(seq, positions=sorted(pos)).cpu().numpy()
    sc.close()
    err = np.abs(got - want)
    nsites = df["mutant"].str.count(":").to_numpy() + 1
    terr = np.abs(tab[pos] - g["table"]).max()
    print(f"\nESM2-3B true size {mode} ({prec}): max|dscore|={err.max():.2e} (1 site {err[nsites == 1].max():.2e}, 5 sites "
          f"{err[nsites == 5].max():.2e}) mean={err.mean():.2e} spearman={spearman(got, want):.6f} max|dlogp|={terr:.2e}")
    assert spearman(got, want) >= 0.999
    if mode == "auto":
        assert err.max() < TOL and terr < 5e-4
    else:
        assert err[nsites == 1].max() < TOL and err.max() < 3e-3 and terr < TOL


def test_multi_site_error_growth_650m_vs_oracle():
    """How the per-mutant error grows with the number of mutated sites at ESM-1v 650M (true size, 96-residue protein, fp32 CPU oracle):
    the measurement behind esm_engine.choose_precision. `auto` must meet 1e-3 for every depth; f16f8 must meet it up to two sites."""
    from proteingym_b200.esm_engine import choose_precision
    arch = synth.ESM1V_650M
    st = synth.make_esm_state(arch, seed=0)
    seq = synth.random_protein(96, 7)
    rng = np.random.RandomState(5)
    muts = {}
    for k in (1, 2, 3, 5):
        lst = []
        for _ in range(150):
            ps = sorted(rng.choice(96, size=k, replace=False))
            lst.append(":".join(f"{seq[p]}{p + 1}{rng.choice([a for a in synth.AA20 if a != seq[p]])}" for p in ps))
        muts[k] = lst
    table = O.masked_marginal_table(O.load_state(st, "esm1v"), seq, "esm1v", arch.layers, arch.heads, batch=16)
    cfg = checkpoint.config_from_synth(arch)
    worst = {}
    for prec in ("f16f8", "f16x3", "f16d"):
        sc = scorer(arch, st, precision=prec, max_rows=65536)
        for k, lst in muts.items():
            got = sc.score_assay(seq, lst).astype(np.float64)
            worst[(prec, k)] = float(np.abs(got - O.score_mutants(lst, seq, table)).max())
        sc.close()
    print("\nmax |score - fp32 oracle| by sites, ESM-1v 650M: " + ", ".join(f"{p} k={k}: {v:.2e}" for (p, k), v in worst.items()))
    for k, lst in muts.items():
        assert worst[(choose_precision(cfg, lst), k)] < TOL
    assert worst[("f16f8", 1)] < TOL and worst[("f16f8", 2)] < TOL
    # delta operands: the error follows the size of the perturbation one mask causes, ~1/L — at 96 residues it sits just inside the bar
    # (7.6e-4 / 8.8e-4 for 1 / 2 sites), which is why the auto rule asks for >= 192 residues before it picks f16d; here only sanity
    assert worst[("f16d", 1)] < 2 * TOL and worst[("f16d", 2)] < 2 * TOL
    assert all(choose_precision(cfg, lst, seq_len=96) != "f16d" for lst in muts.values())
    assert choose_precision(cfg, muts[1], seq_len=286) == "f16d"  # BLAT-sized: measured 3.9e-4 (tests/test_gpu_delta.py)


@pytest.mark.parametrize("L", [1, 2, 1022, 1023])
def test_boundary_lengths_vs_oracle(L):
    """L + 2 = 3, 4 (tiny), 1024 (largest un-windowed input) and 1025 (first windowed one: every row's window drops one token)."""
    arch = synth.EsmArch("esm1v", 1, 64, 1, 64)
    st = synth.make_esm_state(arch, seed=12)
    seq = synth.random_protein(L, L)
    pos = sorted({1, max(1, L // 2), L})
    sc = scorer(arch, st, max_rows=8192)
    got = sc.masked_marginal_table(seq, positions=pos).cpu().numpy()
    sc.close()
    ref = O.masked_marginal_table(O.load_state(st, "esm1v", torch.float64), seq, "esm1v", 1, 1, dtype=torch.float64, positions=pos)
    assert np.abs(got[pos] - ref[pos].numpy()).max() < 2e-4


def test_position_partition_two_gpus_bit_identical():
    """Secondary partitioning: one assay's masked positions split over 2 GPUs + one all-gather of the rows gives,
    bit for bit, the single-GPU scores. Needs 2 devices (skipped on the 1-GPU box; runs under `gpurun --gpus 2`)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29671", os.path.join(root, "scripts", "check_position_partition.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit_identical=True" in r.stdout


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_model_object_seam_reproduces_reference_loop(name, tmp_path):
    """Seam B2 (proteingym_b200.pretrained): the reference's masked-marginal loop (compute_fitness.py:486-504), verbatim, over the
    B200 model object -> the reference's own token_probs table."""
    from proteingym_b200 import pretrained
    g = load_golden(name)
    arch, seq = g["arch"], g["seq"]
    stem = "esm2_tiny" if arch.kind == "esm2" else "esm1v_tiny"
    path = str(tmp_path / f"{stem}.pt")
    synth.write_esm_checkpoint(path, arch, seed=g["meta"]["seed"])
    model, alphabet = pretrained.load_model_and_alphabet(path)
    model.eval()
    model = model.cuda()
    batch_converter = alphabet.get_batch_converter()
    _, _, batch_tokens = batch_converter([("protein1", seq)])
    all_token_probs = []
    for i in range(batch_tokens.size(1)):
        batch_tokens_masked = batch_tokens.clone()
        batch_tokens_masked[0, i] = alphabet.mask_idx
        token_probs = torch.log_softmax(model(batch_tokens_masked.cuda())["logits"], dim=-1)
        all_token_probs.append(token_probs[:, i])
    table = torch.cat(all_token_probs, dim=0).cpu().numpy()
    assert table.shape == g["table"].shape and np.abs(table - g["table"]).max() < 2e-4
