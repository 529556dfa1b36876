"""ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY. Build-container script (needs /root/reference).

Golden vectors for the Tranception path from a hybrid of the reference (see oracle/ref_shims_tranception.py): the reference's
own TranceptionBlock modules + get_slopes inside a thin wrapper (embedding, ALiBi buffer, ln_f, tied lm_head — restating
model_pytorch.py:368-380,499-507,526-612,783), scored by the reference's UNMODIFIED ``score_mutants`` /
``scoring_utils.get_sequence_slices`` / ``get_tranception_scores_mutated_sequences`` (called as unbound functions on the wrapper).
Writes tests/golden/tranception_*.  Checkpoints are regenerated from seeds by proteingym_b200.synth.make_tranception_state."""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

from oracle import ref_shims_tranception as R  # noqa: E402
from proteingym_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_wrapper(mp, arch, st, scoring_window):
    cfg = SimpleNamespace(hidden_size=arch.embed_dim, num_attention_heads=arch.heads, max_position_embeddings=arch.n_ctx,
                          scale_attn_weights=True, attn_pdrop=0.0, resid_pdrop=0.0, attention_mode="tranception", n_inner=arch.ffn_dim,
                          layer_norm_epsilon=arch.ln_eps, add_cross_attention=False, activation_function="squared_relu")

    class Hybrid(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.h = torch.nn.ModuleList([mp.TranceptionBlock(cfg) for _ in range(arch.layers)])
            self.wte = torch.nn.Embedding(arch.vocab, arch.embed_dim)
            self.ln_f = torch.nn.LayerNorm(arch.embed_dim, eps=arch.ln_eps)
            self.lm_head = torch.nn.Linear(arch.embed_dim, arch.vocab, bias=False)
            slopes = torch.Tensor(mp.get_slopes(arch.heads, mode="grouped_alibi"))
            alibi = slopes.unsqueeze(1).unsqueeze(1) * torch.arange(arch.n_ctx).unsqueeze(0).unsqueeze(0).expand(arch.heads, -1, -1)
            self.register_buffer("alibi", alibi.view(arch.heads, 1, arch.n_ctx))
            self.config = SimpleNamespace(tokenizer=R.tokenizer(), scoring_window=scoring_window, retrieval_aggregation_mode=None,
                                          n_ctx=arch.n_ctx)

        @property
        def device(self):
            return torch.device("cpu")

        def encode_batch(self, protein_sequence, sequence_name="sliced_mutated_sequence"):
            return mp.TranceptionLMHeadModel.encode_batch(self, protein_sequence, sequence_name)

        def forward(self, input_ids=None, attention_mask=None, labels=None, return_dict=True, **kw):
            x = self.wte(input_ids)
            am = None
            if attention_mask is not None:
                am = (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * -10000.0
            for blk in self.h:
                x = blk(x, attention_mask=am, alibi_bias=self.alibi)[0]
            return SimpleNamespace(logits=self.lm_head(self.ln_f(x)))

    m = Hybrid()
    sd = {}
    for k, v in st.items():
        if k.startswith("transformer.h."):
            sd[k[len("transformer."):]] = v
        elif k.startswith("transformer."):
            sd[k[len("transformer."):]] = v
        else:
            sd[k] = v
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(("attn.bias" in k or "masked_bias" in k or k == "alibi") for k in missing), missing
    return m.eval()


def run_case(name, arch, seed, target_seq, DMS, indel_mode=False, scoring_window="optimal"):
    mp, su = R.install()
    st = synth.make_tranception_state(arch, seed)
    model = build_wrapper(mp, arch, st, scoring_window)
    with torch.no_grad():
        scores = mp.TranceptionLMHeadModel.score_mutants(model, DMS_data=DMS, target_seq=target_seq, scoring_mirror=True,
                                                          batch_size_inference=20, num_workers=0, indel_mode=indel_mode)
        toks = model.config.tokenizer([target_seq[:min(len(target_seq), 60)], target_seq[:23]], add_special_tokens=True, padding=True,
                                      return_tensors="pt")
        logits = model(**toks).logits
    scores.to_csv(os.path.join(GOLD, f"{name}_reference_scores.csv"), index=False)
    DMS.to_csv(os.path.join(GOLD, f"{name}_dms.csv"), index=False)
    np.save(os.path.join(GOLD, f"{name}_padded_batch_logits.npy"), logits.numpy().astype(np.float32))
    with open(os.path.join(GOLD, f"{name}_meta.json"), "w") as fh:
        json.dump({"name": name, "arch": arch.__dict__, "seed": seed, "target_seq": target_seq, "indel_mode": indel_mode,
                   "scoring_window": scoring_window, "torch": torch.__version__}, fh, indent=1)
    print(f"[gen_golden_tranception] {name}: {len(scores)} scored rows", flush=True)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    arch = synth.TranceptionArch(2, 256, 4, 512)
    seq = synth.random_protein(60, 31)
    muts = synth.sample_mutants(seq, 120, seed=3, multi_frac=0.3) + [None]
    df = pd.DataFrame({"mutant": [m for m in muts if m], "DMS_score": 0.0})
    df["mutated_sequence"] = [synth.apply_mutant(seq, m) for m in df["mutant"]]
    run_case("tranception_subs", arch, 5, seq, df)
    seq2 = synth.random_protein(45, 32)
    ind = synth.random_indels(seq2, 50, seed=4) + [seq2]
    run_case("tranception_indels", arch, 5, seq2, pd.DataFrame({"mutant": ind, "mutated_sequence": ind, "DMS_score": 0.0}), indel_mode=True)
    seq3 = synth.random_protein(1100, 33)   # longer than n_ctx - 2 = 1022: optimal windows around the mutation barycentre
    muts3 = synth.sample_mutants(seq3, 40, seed=6, multi_frac=0.25)
    df3 = pd.DataFrame({"mutant": muts3, "DMS_score": 0.0})
    df3["mutated_sequence"] = [synth.apply_mutant(seq3, m) for m in muts3]
    run_case("tranception_long", synth.TranceptionArch(1, 256, 4, 256), 7, seq3, df3)


if __name__ == "__main__":
    main()
