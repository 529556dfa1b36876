"""EVE log-prior pre-step at the reference's real sizes (encoder 2000-1000-300, z 50, decoder 300-1000-2000, 40-channel output
convolution; utils/eve_model_default_params.json) for a 500-column alignment: the reference-order sampler (weights drawn one by
one, timed on a few samples) against the batched local-reparameterisation sampler at the launcher's 200 000 samples. One JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proteingym_b200 import eve_prior, synth  # noqa: E402

PARAMS = {
    "encoder_parameters": {"hidden_layers_sizes": [2000, 1000, 300], "z_dim": 50, "convolve_input": False, "convolution_input_depth": 40,
                           "nonlinear_activation": "relu", "dropout_proba": 0.0},
    "decoder_parameters": {"hidden_layers_sizes": [300, 1000, 2000], "z_dim": 50, "bayesian_decoder": True, "first_hidden_nonlinearity": "relu",
                           "last_hidden_nonlinearity": "relu", "dropout_proba": 0.1, "convolve_output": True, "convolution_output_depth": 40,
                           "include_temperature_scaler": True, "include_sparsity": False, "num_tiles_sparsity": 0, "logit_sparsity_p": 0},
}


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    st = synth.make_eve_state(L, PARAMS, seed=1, log_var=-6.0)
    focus, cols = list(synth.random_protein(L, 5)), list(range(L))

    def timed(n, how):
        if dev == "cuda":
            torch.cuda.synchronize()
        t = time.time()
        out = eve_prior.eve_log_prior_single(st, PARAMS, focus, cols, L, 0, n, device=dev, sampler=how)
        if dev == "cuda":
            torch.cuda.synchronize()
        return time.time() - t, out

    timed(2, "stream"); timed(64, "local")  # warm-up (cuBLAS handles, allocator)
    n_stream = 200 if dev == "cuda" else 3
    ts, a = timed(n_stream, "stream")
    tl, b = timed(N, "local")
    fin = torch.isfinite(a)
    print(json.dumps({"L": L, "device": dev, "stream_samples": n_stream, "stream_s": round(ts, 3), "stream_ms_per_sample": round(1e3 * ts / n_stream, 3),
                      "stream_extrapolated_s_at_N": round(ts / n_stream * N, 1), "local_samples": N, "local_s": round(tl, 3),
                      "local_us_per_sample": round(1e6 * tl / N, 2), "speedup_at_N": round(ts / n_stream * N / tl, 1),
                      "mean_abs_diff_stream_vs_local": round((a[fin] - b[fin]).abs().mean().item(), 4)}))


if __name__ == "__main__":
    main()
