"""GPU (-m gpu): the delta-operand mode (PG_PREC_F16D) — kernel-level checks of the three epilogue features it adds (GEMM base rows
and masked-row skip, attention base subtraction and compact exact rows) against fp64, then the mode end to end against the oracle
and the unmodified reference's golden CSV at true ESM-1v 650M size (tests/golden/blat_esm1v_650m). Tolerance: 1e-3 abs per mutant."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import esm_oracle as O
from proteingym_b200 import _lib, synth
from test_gpu_parity import TOL, hilo, scorer, spearman

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M_copies,T,N,K,epi", [(3, 50, 128, 64, 0), (2, 130, 320, 128, 1), (4, 77, 192, 256, 2), (2, 514, 1280, 1280, 2),
                                                (1, 300, 256, 128, 0)])
def test_gemm_delta_epilogue_matches_fp64(M_copies, T, N, K, epi):
    """C = epi(base_pre[t] + D W^T) [- base_post[t]], masked rows untouched with the residual epilogue."""
    lib = _lib.load()
    M = M_copies * T
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    D = (torch.randn(M, K, device="cuda", generator=g) * 0.1).half().contiguous()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half().contiguous()
    base_pre = torch.randn(T, N, device="cuda", generator=g)
    base_post = torch.randn(T, N, device="cuda", generator=g).half()   # the activated base row is an fp16 plane
    mask_pos = torch.randint(0, T, (M_copies,), device="cuda", generator=g, dtype=torch.int32)
    t = torch.arange(M, device="cuda") % T
    acc = D.double() @ W.double().T + base_pre.double()[t]
    args = _lib.PgGemmArgs()
    args.a, args.lda, args.w, args.ldw, args.bias = D.data_ptr(), K, W.data_ptr(), K, None
    args.M, args.N, args.K, args.nseg, args.epi = M, N, K, 1, epi
    args.base_pre, args.base_T = base_pre.data_ptr(), T
    if epi == 2:
        resid = torch.randn(M, N, device="cuda", generator=g)
        r0 = resid.clone()
        args.resid, args.ldr, args.mask_pos = resid.data_ptr(), N, mask_pos.data_ptr()
        ref = r0.double() + acc
        masked = torch.arange(M_copies, device="cuda") * T + mask_pos.long()
        ref[masked] = r0.double()[masked]
    else:
        out = torch.zeros(M, 2 * N, device="cuda", dtype=torch.float16)
        args.out_h, args.ldo, args.out_lo_off, args.out_fmt = out.data_ptr(), 2 * N, N, 1
        if epi == 1:
            acc = acc * 0.5 * (1 + torch.erf(acc / 2 ** 0.5))
            args.base_post = base_post.data_ptr()
            acc = acc - base_post.double()[t]
        ref = acc
    _lib.check(lib.pg_gemm(C.byref(args), None))
    torch.cuda.synchronize()
    got = resid.double() if epi == 2 else out[:, :N].double() + out[:, N:].double()
    err = (got - ref).abs().max().item()
    assert err < 4e-6 * max(4.0, ref.abs().max().item()), err
    if epi == 2:
        assert torch.equal(resid[masked], r0[masked])  # bit-untouched


@pytest.mark.parametrize("B,T,H", [(3, 70, 2), (2, 514, 4), (5, 130, 1)])
def test_attention_delta_outputs_match_fp64(B, T, H):
    """out = attention - base_o[t] as a single fp16 plane; exact hi/lo copies of the rows mask_pos[b]."""
    lib = _lib.load()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * T + H)
    qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g)
    qkv[:, :d] *= 0.3
    q16 = hilo(qkv)
    eff = q16[:, :3 * d].double() + q16[:, 3 * d:].double()
    q, k, v = [eff[:, i * d:(i + 1) * d].view(B, T, H, 64).transpose(1, 2) for i in range(3)]
    ref = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(B * T, d)
    base = (ref.view(B, T, d)[0].float() + 0.05 * torch.randn(T, d, device="cuda", generator=g)).half().contiguous()  # fp16 base rows
    mask_pos = torch.randint(0, T, (B,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.zeros(B * T, 2 * d, device="cuda", dtype=torch.float16)
    cout = torch.zeros(B, 2 * d, device="cuda", dtype=torch.float16)
    a = _lib.PgAttnArgs()
    a.qkv, a.ld, a.lo_off = q16.data_ptr(), 6 * d, 3 * d
    a.out, a.ldo, a.out_lo_off, a.out_fmt = out.data_ptr(), 2 * d, 0, 0
    a.B, a.T, a.heads, a.nseg, a.causal, a.impl = B, T, H, 3, 0, 0
    a.base_o, a.mask_pos, a.cout, a.ldc, a.c_lo_off = base.data_ptr(), mask_pos.data_ptr(), cout.data_ptr(), 2 * d, d
    _lib.check(lib.pg_attention(C.byref(a), None))
    torch.cuda.synchronize()
    t = torch.arange(B * T, device="cuda") % T
    want = ref - base.double()[t]
    assert (out[:, :d].double() - want).abs().max().item() < 2 ** -11 * max(1.0, want.abs().max().item()) + 1e-5  # one fp16 plane
    assert torch.count_nonzero(out[:, d:]) == 0
    rows = torch.arange(B, device="cuda") * T + mask_pos.long()
    assert (cout[:, :d].double() + cout[:, d:].double() - ref[rows]).abs().max().item() < 2e-5


@pytest.mark.parametrize("L,layers,d,heads,ffn,lnb", [(70, 2, 128, 2, 256, False), (130, 2, 128, 2, 256, True), (37, 1, 64, 1, 64, False),
                                                      (257, 4, 256, 4, 1024, False)])
def test_delta_mode_matches_oracle(L, layers, d, heads, ffn, lnb):
    arch = synth.EsmArch("esm1v", layers, d, heads, ffn, emb_layer_norm_before=lnb)
    st = synth.make_esm_state(arch, seed=3)
    seq = synth.random_protein(L, 11)
    ref = O.masked_marginal_table(O.load_state(st, "esm1v", torch.float64), seq, "esm1v", layers, heads, dtype=torch.float64,
                                  positions=range(1, L + 1))
    muts = synth.sample_mutants(seq, 300, 5, multi_frac=0.3)
    want = O.score_mutants(muts, seq, ref)
    for max_rows in (32768, 4 * (L + 2)):  # one pass, then several passes of four copies: same scores
        sc = scorer(arch, st, precision="f16d", max_rows=max_rows)
        table = sc.masked_marginal_table(seq).cpu().double()
        assert (table[1:L + 1] - ref[1:L + 1]).abs().max().item() < 3e-4
        got = sc.score_assay(seq, muts).astype(np.float64)
        assert np.abs(got - want).max() < TOL
        sc.close()


def test_delta_mode_rejects_other_architectures_and_falls_back_for_windows():
    from proteingym_b200 import checkpoint
    from proteingym_b200.esm_engine import EsmScorer
    arch2 = synth.EsmArch("esm2", 2, 128, 2, 256)
    with pytest.raises(Exception):
        EsmScorer(checkpoint.config_from_synth(arch2), checkpoint.normalise_synth_state(arch2, synth.make_esm_state(arch2, seed=1)), precision="f16d")
    # a protein longer than the model window: per-position windows differ, the pass runs at fp16 hi/lo x3 precision instead
    arch = synth.EsmArch("esm1v", 1, 64, 1, 64)
    st = synth.make_esm_state(arch, seed=2)
    seq = synth.random_protein(1100, 5)
    pos = [1, 400, 700, 1100]
    a = scorer(arch, st, precision="f16d", max_rows=8192).masked_marginal_table(seq, positions=pos)
    b = scorer(arch, st, precision="f16x3", max_rows=8192).masked_marginal_table(seq, positions=pos)
    assert torch.equal(a[pos], b[pos])


def test_delta_mode_golden_true_size_blat():
    """BASELINE.json config 1 at true ESM-1v 650M size against the unmodified reference CLI's CSV (fp32 CPU)."""
    g = load_golden("blat_esm1v_650m")
    arch, seq, df = g["arch"], g["seq"], g["df"]
    col = g["meta"]["ckpt_names"][0].split(".")[0]
    want = df[col].to_numpy()
    sc = scorer(arch, g["state"](), precision="f16d", max_rows=131072)
    got = sc.score_assay(seq, list(df["mutant"]))
    err = np.abs(got - want)
    tab = sc.masked_marginal_table(seq, positions=range(24, 287)).cpu().numpy()
    terr = np.abs(tab[24:287] - g["table"][24:287]).max()
    print(f"\nBLAT 650M f16d: max|d|={err.max():.2e} mean={err.mean():.2e} spearman={spearman(got, want):.6f} max|dlogp|={terr:.2e}")
    assert err.max() < TOL and spearman(got, want) >= 0.999
    sc.close()
