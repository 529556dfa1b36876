"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

CPU restatement (plain torch tensor ops, fp32 or fp64, no nn.Module) of ProteinGym's ESM-1b/ESM-1v/ESM2
masked-marginal scoring path. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; ``proteingym_b200`` never does.

Parity status: PINNED against the reference itself. ``oracle/gen_golden.py`` runs the UNMODIFIED reference
(``/root/reference/proteingym/baselines/esm/compute_fitness.py::main`` and the vendored fair-esm modules) in the build
container on seeded synthetic checkpoints and commits its outputs under ``tests/golden/``;
``tests/test_oracle_vs_golden.py`` checks this restatement against those vectors (the reference's own test-suite
holds no vectors for this path).

Every function cites the reference lines it restates (paths relative to /root/reference/proteingym/baselines/esm).
"""
from __future__ import annotations

import math

import numpy as np
import torch

MASK_IDX = 32
PAD_IDX = 1
CLS_IDX = 0
EOS_IDX = 2
VOCAB = ['<cls>', '<pad>', '<eos>', '<unk>', 'L', 'A', 'G', 'V', 'S', 'E', 'R', 'T', 'I', 'D', 'P', 'K', 'Q', 'N',
         'F', 'Y', 'M', 'H', 'W', 'C', 'X', 'B', 'U', 'Z', 'O', '.', '-', '<null_1>', '<mask>']
TOK = {t: i for i, t in enumerate(VOCAB)}


def tokenize(seq: str) -> torch.Tensor:
    """esm/data.py:262-297 (BatchConverter, batch of one, prepend_bos & append_eos)."""
    return torch.tensor([CLS_IDX] + [TOK[c] for c in seq] + [EOS_IDX], dtype=torch.int64)


def gelu(x):
    """esm/modules.py:17-24 (exact erf GELU)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps=1e-5):
    """esm/modules.py:68-81 (ESM1bLayerNorm == torch.nn.LayerNorm, eps 1e-5, biased variance)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def rotate_half(x):
    """esm/rotary_embedding.py:11-13."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def rotary_tables(T, hd, dtype):
    """esm/rotary_embedding.py:37-61: inv_freq computed in fp32, tables over token index 0..T-1."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
    t = torch.arange(T).type_as(inv_freq)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def load_state(st: dict, kind: str, dtype=torch.float32) -> dict:
    """Checkpoint key handling of esm/pretrained.py:85-99 (v1) / :162-181 (v2), restated.

    ``st`` holds un-prefixed keys. v1 zeroes the ``<mask>`` row of ``embed_tokens.weight`` in place (:97). The tied
    ``lm_head.weight`` entry is copied last by ``load_state_dict`` into the shared parameter (module registration
    order, esm/model/esm1.py:66-102), so it wins; in released checkpoints (and ours) both keys alias one storage, so
    the zeroing is visible through either name."""
    emb = st["embed_tokens.weight"]
    head = st.get("lm_head.weight", emb)
    aliased = head.data_ptr() == emb.data_ptr()
    emb = emb.clone()
    if kind == "esm1v":
        emb[MASK_IDX].zero_()
    head = emb if aliased else head.clone()
    out = {k: v.to(dtype) for k, v in st.items() if k not in ("embed_tokens.weight", "lm_head.weight")}
    out["embed_tokens.weight"] = head.to(dtype)
    out["lm_head.weight"] = head.to(dtype)
    return out


def esm_forward(st: dict, tokens: torch.Tensor, kind: str, layers: int, heads: int, token_dropout=True,
                dtype=torch.float32, rnd=None, mm=None) -> torch.Tensor:
    """logits [B,T,V]. Restates ProteinBertModel.forward (esm/model/esm1.py:116-193, ESM-1b branch) and
    ESM2.forward (esm/model/esm2.py:76-143); TransformerLayer (esm/modules.py:120-142); MultiheadAttention manual
    path (esm/multihead_attention.py:242-395; the F.multi_head_attention_forward fast path :196-230 is the same
    arithmetic); RobertaLMHead (esm/modules.py:322-328). No padding is ever present on this path (batch of one
    sequence or equal-length windows), so padding-mask branches reduce to identity.

    ``rnd`` (default None = exact) is a numerics-study hook, not reference behaviour: a function applied to every
    tensor-core GEMM operand (e.g. round-trip through fp16) to emulate the CUDA path's operand rounding on CPU.
    ``mm`` (default None = exact ``a @ w.T``) is the same kind of hook for the four linear layers as a whole: ``mm(a, w, site)``
    with site in {"qkv", "out", "fc1", "fc2"} returns the emulated product (used to study split-operand schemes whose result is
    not a product of two rounded operands, e.g. fp16 hi*hi + fp8 cross terms)."""
    if rnd is None:
        rnd = lambda t: t
    if mm is None:
        mm = lambda a, w, site: rnd(a) @ rnd(w).T
    E = st["lm_head.weight"].to(dtype)  # shared embedding / output matrix, see load_state
    B, T = tokens.shape
    d = E.shape[1]
    hd = d // heads
    x = E[tokens]  # embed_scale == 1 (esm1.py:90, esm2.py:41)
    if token_dropout:
        is_mask = tokens == MASK_IDX
        x = x.masked_fill(is_mask.unsqueeze(-1), 0.0)  # esm1.py:126 / esm2.py:86
        src_lengths = torch.full((B,), T, dtype=dtype)
        ratio = is_mask.sum(-1).to(dtype) / src_lengths
        x = x * (1 - 0.15 * 0.8) / (1 - ratio)[:, None, None]  # esm1.py:128-131
    if kind == "esm1v":
        pos = torch.arange(T) + PAD_IDX + 1  # esm/modules.py:263-264 with no padding: cumsum(1..T) + 1
        x = x + st["embed_positions.weight"].to(dtype)[pos][None]
        if "emb_layer_norm_before.weight" in st:
            x = layer_norm(x, st["emb_layer_norm_before.weight"].to(dtype), st["emb_layer_norm_before.bias"].to(dtype))
    else:
        cos, sin = rotary_tables(T, hd, dtype)
    for i in range(layers):
        p = f"layers.{i}."
        W = lambda n: st[p + n].to(dtype)
        h = layer_norm(x, W("self_attn_layer_norm.weight"), W("self_attn_layer_norm.bias"))
        q = mm(h, W("self_attn.q_proj.weight"), "qkv") + W("self_attn.q_proj.bias")
        k = mm(h, W("self_attn.k_proj.weight"), "qkv") + W("self_attn.k_proj.bias")
        v = mm(h, W("self_attn.v_proj.weight"), "qkv") + W("self_attn.v_proj.bias")
        q = q * hd ** -0.5  # multihead_attention.py:261
        q = q.view(B, T, heads, hd).transpose(1, 2)
        k = k.view(B, T, heads, hd).transpose(1, 2)
        v = v.view(B, T, heads, hd).transpose(1, 2)
        if kind == "esm2":
            q = q * cos + rotate_half(q) * sin  # rotary_embedding.py:16-20, after scaling (:261 then :354)
            k = k * cos + rotate_half(k) * sin
        q, k, v = rnd(q), rnd(k), rnd(v)
        a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)  # :357, :379
        o = (rnd(a) @ v).transpose(1, 2).reshape(B, T, d)  # :387-394
        x = x + mm(o, W("self_attn.out_proj.weight"), "out") + W("self_attn.out_proj.bias")
        h = layer_norm(x, W("final_layer_norm.weight"), W("final_layer_norm.bias"))
        h = gelu(mm(h, W("fc1.weight"), "fc1") + W("fc1.bias"))
        x = x + mm(h, W("fc2.weight"), "fc2") + W("fc2.bias")
    x = layer_norm(x, st["emb_layer_norm_after.weight"].to(dtype), st["emb_layer_norm_after.bias"].to(dtype))
    h = gelu(x @ st["lm_head.dense.weight"].to(dtype).T + st["lm_head.dense.bias"].to(dtype))
    h = layer_norm(h, st["lm_head.layer_norm.weight"].to(dtype), st["lm_head.layer_norm.bias"].to(dtype))
    return h @ E.T + st["lm_head.bias"].to(dtype)


def get_optimal_window(mutation_position_relative, seq_len_wo_special, model_window):
    """proteingym/utils/scoring_utils.py:43-52."""
    half_model_window = model_window // 2
    if seq_len_wo_special <= model_window:
        return [0, seq_len_wo_special]
    elif mutation_position_relative < half_model_window:
        return [0, model_window]
    elif mutation_position_relative >= seq_len_wo_special - half_model_window:
        return [seq_len_wo_special - model_window, seq_len_wo_special]
    else:
        return [max(0, mutation_position_relative - half_model_window),
                min(seq_len_wo_special, mutation_position_relative + half_model_window)]


def masked_marginal_table(st, seq, kind, layers, heads, token_dropout=True, dtype=torch.float32, positions=None,
                          batch=8):
    """compute_fitness.py:486-504: one masked copy per token index i in 0..L+1 (BOS/EOS included), optimal 1024-window
    when L+2 > 1024 (:492-495), keep ``log_softmax(logits)[i-start]``. Returns [L+2, 33]; rows not in ``positions``
    (if given) are left as NaN. Copies are batched (``batch``) — equal length, no padding, so identical arithmetic."""
    tokens = tokenize(seq)
    T = tokens.numel()
    idx = list(range(T)) if positions is None else list(positions)
    table = torch.full((T, 33), float("nan"), dtype=dtype)
    # group by window start so batched rows have equal length/content apart from the mask
    groups = {}
    for i in idx:
        start, end = (get_optimal_window(i, T, 1024) if T > 1024 else (0, T))
        groups.setdefault((start, end), []).append(i)
    for (start, end), members in groups.items():
        for c in range(0, len(members), batch):
            chunk = members[c:c + batch]
            tb = tokens[start:end].unsqueeze(0).repeat(len(chunk), 1)
            for r, i in enumerate(chunk):
                tb[r, i - start] = MASK_IDX
            lp = torch.log_softmax(esm_forward(st, tb, kind, layers, heads, token_dropout, dtype), dim=-1)
            for r, i in enumerate(chunk):
                table[i] = lp[r, i - start]
    return table


def label_row(row: str, sequence: str, table: torch.Tensor, offset_idx: int = 1) -> float:
    """compute_fitness.py:240-250 (sum over ':'-separated sites of table[1+idx, mt] - table[1+idx, wt])."""
    score = 0
    for mutation in row.split(":"):
        wt, idx, mt = mutation[0], int(mutation[1:-1]) - offset_idx, mutation[-1]
        assert sequence[idx] == wt, "The listed wildtype does not match the provided sequence"
        score += (table[1 + idx, TOK.get(mt, 3)] - table[1 + idx, TOK.get(wt, 3)]).item()
    return score


def score_mutants(mutants, sequence, table, offset_idx=1) -> np.ndarray:
    return np.array([label_row(m, sequence, table, offset_idx) for m in mutants], dtype=np.float64)
