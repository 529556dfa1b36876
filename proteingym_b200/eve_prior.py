"""EVE log prior for TranceptEVE (trancepteve/model_pytorch.py:940-1001): the once-per-protein pre-step that turns trained EVE
VAE checkpoints into an ``EVE_log_prior`` [L_full, 25] table (-inf outside the MSA's focus columns and on the 5 special tokens).

This is not part of the scoring hot path and is cached on disk exactly where the reference caches it
(``<EVE model folder>/log_prior/<model name>_<num samples>_log_space``, a pickled tensor), so a cache written by either
implementation is read by the other. When there is no cache the Monte-Carlo average over the Bayesian decoder is computed here
with plain tensor algebra on the scorer's device, drawing from a ``torch.Generator`` seeded like the reference (VAE_model.py:37
``torch.manual_seed(random_seed)``, seed 42) in the reference's draw order (latent, then per decoder layer weight -> bias,
output weight -> bias, output convolution, [sparsity], temperature), so that with the same torch build and device type the samples
are the same numbers. torch's generator is used on purpose: stream-for-stream agreement with the reference's sampling is only
possible through it; the arithmetic is restated, not imported.

For the sample counts the reference's launcher asks for (``EVE_num_samples_log_proba=200000``,
scripts/scoring_DMS_zero_shot/scoring_TranceptEVE_substitutions.sh) drawing every decoder weight per sample is 40 M normals and
320 MB of parameter traffic per sample for a 500-residue protein. ``sampler="local"`` draws from the SAME distribution without
materialising weights (local reparameterisation): with independent Gaussian weight posteriors, a Bayesian linear layer's
pre-activations given its input h are independent Gaussians N(mu_W h + mu_b, sigma_W^2 (h*h) + sigma_b^2), and the output layer's
double reinterpretation + 1x1 convolution reduces (when alphabet | last hidden size) to one dot product per logit between a
contiguous 1/alphabet slice of the output weights and y[j, c] = sum_a h[j*A + a] conv[c, a]. Thousands of samples then go through a
handful of dense GEMMs per batch. ``sampler="auto"`` keeps the stream-exact path up to 2000 samples and switches to ``local`` above.

Kernels. On a CUDA device every dense product of this file — encoder layers, the decoder's Bayesian layers (means and variances),
the output layer and the 1x1 output convolution — runs on the library's own kernels, not on torch / cuBLAS: ``_mm`` packs both
operands into fp16 hi/lo rows (``pg_pack_weight`` fmt 1, after an exact power-of-two scaling into fp16's comfortable range) and calls
``pg_gemm`` (tcgen05 GEMM, three hi/lo passes = ~2^-22 relative per product, fp32 reduce-add epilogue); the per-sample output
convolution of the batched sampler is ``pg_eve_output_conv``. Packed weight matrices are cached for the duration of one
``eve_log_prior_single`` call (the batched sampler reuses them for ~100 batches). torch supplies only the random streams (see
above), element-wise glue and ``log_softmax``. With CPU tensors (the CPU test-suite, which pins this file's arithmetic against the
reference's functions) the same code path uses ``torch.matmul`` — there is no CUDA device to run a kernel on; the product
(``TranceptEVEScorer``) always passes its GPU.

State-dict keys / shapes: VAE_encoder.py:40-52, VAE_decoder.py:47-108."""
from __future__ import annotations

import ctypes
import json
import math
import os
import pickle

import numpy as np
import torch

ALPHABET = "ACDEFGHIKLMNPQRSTVWY"


class _Gemm:
    """``x @ w.T`` for 2-D fp32 tensors on the library's tcgen05 GEMM (module docstring). ``key``: cache the packed form of ``w``
    under that name (weights that do not change between calls)."""

    def __init__(self):
        self.cache = {}
        self.lib = None

    @staticmethod
    def _pow2(t: torch.Tensor) -> float:
        m = float(t.abs().max())
        return 1.0 if not (m > 0.0 and math.isfinite(m)) else 2.0 ** (9 - math.floor(math.log2(m)))  # largest |value| into [512, 1024)

    def _pack(self, t: torch.Tensor):
        from . import _lib
        R, K = t.shape
        Kp = (K + 63) // 64 * 64
        sc = self._pow2(t)
        src = torch.zeros((R, Kp), dtype=torch.float32, device=t.device)
        src[:, :K] = t * sc
        out = torch.empty((R, 2 * Kp), dtype=torch.float16, device=t.device)
        _lib.check(self.lib.pg_pack_weight(src.data_ptr(), R, Kp, 1, out.data_ptr(), None, torch.cuda.current_stream(t.device).cuda_stream))
        return out, sc, Kp

    def __call__(self, x: torch.Tensor, w: torch.Tensor, key: str = None) -> torch.Tensor:
        if x.device.type != "cuda":
            return x @ w.T
        from . import _lib
        if self.lib is None:
            self.lib = _lib.load()
        x = x.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        if key is not None and key in self.cache:
            w16, sw, Kp = self.cache[key]
        else:
            w16, sw, Kp = self._pack(w.contiguous().float())
            if key is not None:
                self.cache[key] = (w16, sw, Kp)
        a16, sa, _ = self._pack(x)
        Np = (N + 3) // 4 * 4
        out = torch.zeros((M, Np), dtype=torch.float32, device=x.device)
        g = _lib.PgGemmArgs()
        g.a, g.lda, g.w, g.ldw, g.bias = a16.data_ptr(), 2 * Kp, w16.data_ptr(), 2 * Kp, None
        g.M, g.N, g.K, g.nseg, g.epi = M, N, Kp, 3, 2
        g.resid, g.ldr = out.data_ptr(), Np
        with torch.cuda.device(x.device):
            _lib.check(self.lib.pg_gemm(ctypes.byref(g), torch.cuda.current_stream(x.device).cuda_stream))
        return out[:, :N] * (1.0 / (sa * sw))


def _output_conv(x: torch.Tensor, conv: torch.Tensor, J: int, A: int, Cd: int) -> torch.Tensor:
    """y[s, j, c] = sum_a x[s, j*A + a] conv[s, c, a] (one convolution draw per sample): ``pg_eve_output_conv`` on the GPU."""
    S = x.shape[0]
    if x.device.type != "cuda":
        return torch.einsum("sja,sca->sjc", x.reshape(S, J, A), conv).reshape(S, J * Cd)
    from . import _lib
    x, conv = x.contiguous().float(), conv.contiguous().float()
    y = torch.empty((S, J * Cd), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pg_eve_output_conv(x.data_ptr(), conv.data_ptr(), S, J, A, Cd, y.data_ptr(),
                                                   torch.cuda.current_stream(x.device).cuda_stream))
    return y
_ACT = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "elu": torch.nn.functional.elu, "linear": lambda x: x}


def encode_focus(st: dict, enc: dict, x: torch.Tensor, mm: _Gemm = None):
    """VAE_MLP_encoder.forward (VAE_encoder.py:66-87): x [B, L, 20] one-hot -> (z_mean, z_log_var)."""
    mm = mm or _Gemm()
    B, L, A = x.shape
    if enc.get("convolve_input"):
        cw = st["encoder.input_convolution.weight"][:, :, 0]                    # [Cin, A]
        x = mm(x.reshape(B * L, A), cw).reshape(B, L, -1).permute(0, 2, 1).reshape(B, -1)   # "bla,ca->bcl"
    else:
        x = x.reshape(B, L * A)
    act = _ACT[enc["nonlinear_activation"]]
    for k in range(len(enc["hidden_layers_sizes"])):
        x = act(mm(x, st[f"encoder.hidden_layers.{k}.weight"]) + st[f"encoder.hidden_layers.{k}.bias"])
    return (mm(x, st["encoder.fc_mean.weight"]) + st["encoder.fc_mean.bias"],
            mm(x, st["encoder.fc_log_var.weight"]) + st["encoder.fc_log_var.bias"])


def _draw(mean: torch.Tensor, log_var: torch.Tensor, g: torch.Generator) -> torch.Tensor:
    eps = torch.randn(mean.shape, generator=g, device=mean.device, dtype=mean.dtype)
    return torch.exp(0.5 * log_var) * eps + mean


def decode_sample(st: dict, dec: dict, z: torch.Tensor, seq_len: int, g: torch.Generator, mm: _Gemm = None) -> torch.Tensor:
    """One pass of VAE_Bayesian_MLP_decoder.forward in eval mode (VAE_decoder.py:118-169) -> log-softmax [B, L, 20]."""
    mm = mm or _Gemm()
    A = len(ALPHABET)
    H = dec["hidden_layers_sizes"]
    x = z
    for k in range(len(H)):
        w = _draw(st[f"decoder.hidden_layers_mean.{k}.weight"], st[f"decoder.hidden_layers_log_var.{k}.weight"], g)
        b = _draw(st[f"decoder.hidden_layers_mean.{k}.bias"], st[f"decoder.hidden_layers_log_var.{k}.bias"], g)
        act = _ACT[dec["first_hidden_nonlinearity"] if k < len(H) - 1 else dec["last_hidden_nonlinearity"]]
        x = act(mm(x, w) + b)
    w_out = _draw(st["decoder.last_hidden_layer_weight_mean"], st["decoder.last_hidden_layer_weight_log_var"], g)
    b_out = _draw(st["decoder.last_hidden_layer_bias_mean"], st["decoder.last_hidden_layer_bias_log_var"], g)
    if dec["convolve_output"]:
        C = dec["convolution_output_depth"]
        conv = _draw(st["decoder.output_convolution_mean.weight"], st["decoder.output_convolution_log_var.weight"], g)
        # the reference reinterprets the buffers (no transposes): [C*L, H] read as [L*H, C], [A, C, 1] read as [C, A]
        w_out = mm(w_out.reshape(seq_len * H[-1], C), conv.reshape(C, A).T)
    if dec.get("include_sparsity"):
        tiles = dec["num_tiles_sparsity"]
        sp = _draw(st["decoder.sparsity_weight_mean"], st["decoder.sparsity_weight_log_var"], g)
        sp = torch.sigmoid(sp.repeat(tiles, 1)).unsqueeze(2)
        w_out = w_out.reshape(H[-1], seq_len, A) * sp
    w_out = w_out.reshape(seq_len * A, H[-1])
    x = mm(x, w_out) + b_out
    if dec["include_temperature_scaler"]:
        t = _draw(st["decoder.temperature_scaler_mean"], st["decoder.temperature_scaler_log_var"], g)
        x = torch.log(1.0 + torch.exp(t)) * x
    return torch.log_softmax(x.reshape(z.shape[0], seq_len, A), dim=-1)


def local_sampling_supported(dec: dict) -> bool:
    return bool(dec["convolve_output"]) and not dec.get("include_sparsity") and dec["hidden_layers_sizes"][-1] % len(ALPHABET) == 0


def decode_batch_local(st: dict, dec: dict, z: torch.Tensor, seq_len: int, g: torch.Generator, mm: _Gemm = None) -> torch.Tensor:
    """``S`` decoder passes at once, each with its own weight draw, by sampling pre-activations instead of weights (see the module
    docstring). z [S, z_dim] -> log-softmax [S, L, 20]. Same distribution as ``decode_sample`` applied to each row of z."""
    mm = mm or _Gemm()
    A = len(ALPHABET)
    H = dec["hidden_layers_sizes"]
    S = z.shape[0]

    def randn(*shape):
        return torch.randn(shape, generator=g, device=z.device, dtype=z.dtype)

    def bayes_linear(x, wm, wlv, bm, blv, key):
        mean = mm(x, wm, key + ".mean") + bm
        var = mm(x * x, torch.exp(wlv), key + ".var") + torch.exp(blv)
        return mean + torch.sqrt(var) * randn(*mean.shape)

    x = z
    for k in range(len(H)):
        act = _ACT[dec["first_hidden_nonlinearity"] if k < len(H) - 1 else dec["last_hidden_nonlinearity"]]
        x = act(bayes_linear(x, st[f"decoder.hidden_layers_mean.{k}.weight"], st[f"decoder.hidden_layers_log_var.{k}.weight"],
                             st[f"decoder.hidden_layers_mean.{k}.bias"], st[f"decoder.hidden_layers_log_var.{k}.bias"], f"hidden{k}"))
    C = dec["convolution_output_depth"]
    J = H[-1] // A
    cm, clv = st["decoder.output_convolution_mean.weight"].reshape(C, A), st["decoder.output_convolution_log_var.weight"].reshape(C, A)
    conv = cm + torch.exp(0.5 * clv) * randn(S, C, A)                      # the small 1x1-convolution weights are drawn directly
    y = _output_conv(x, conv, J, A, C)
    wm = st["decoder.last_hidden_layer_weight_mean"].reshape(seq_len * A, J * C)   # logit q reads the q-th contiguous slice
    wv = torch.exp(st["decoder.last_hidden_layer_weight_log_var"]).reshape(seq_len * A, J * C)
    out = (mm(y, wm, "out.mean") + st["decoder.last_hidden_layer_bias_mean"]) + torch.sqrt(
        mm(y * y, wv, "out.var") + torch.exp(st["decoder.last_hidden_layer_bias_log_var"])) * randn(S, seq_len * A)
    if dec["include_temperature_scaler"]:
        t = st["decoder.temperature_scaler_mean"] + torch.exp(0.5 * st["decoder.temperature_scaler_log_var"]) * randn(S, 1)
        out = torch.log(1.0 + torch.exp(t)) * out
    return torch.log_softmax(out.reshape(S, seq_len, A), dim=-1)


def eve_log_prior_single(state: dict, params: dict, focus_seq_trimmed, focus_cols, full_sequence_len: int, MSA_start: int,
                         num_samples: int = 10, device="cuda", seed: int = 42, sampler: str = "auto", batch: int = 2048) -> torch.Tensor:
    """get_EVE_log_prior_single (model_pytorch.py:969-1001) for the wild type: [full_sequence_len, 25] float32 on ``device``.
    ``sampler``: "stream" (reference's draw order), "local" (same distribution, batched; module docstring) or "auto"."""
    dev = torch.device(device)
    st = {k: v.to(dev, torch.float32) for k, v in state.items()}
    L, A = len(focus_seq_trimmed), len(ALPHABET)
    x = torch.zeros((1, L, A), dtype=torch.float32, device=dev)
    for j, ch in enumerate(focus_seq_trimmed):
        k = ALPHABET.find(ch)
        if k >= 0:
            x[0, j, k] = 1.0
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mm = _Gemm()   # packed-operand cache for this call
    mu, log_var = encode_focus(st, params["encoder_parameters"], x, mm)
    dec = params["decoder_parameters"]
    if sampler == "auto":
        sampler = "local" if num_samples > 2000 and local_sampling_supported(dec) else "stream"
    if sampler == "local":
        if not local_sampling_supported(dec):
            raise ValueError("local sampling needs convolve_output, no sparsity and alphabet | last hidden size")
        recon = torch.zeros((1, L, A), dtype=torch.float64, device=dev)
        done = 0
        while done < num_samples:
            S = min(batch, num_samples - done)
            z = mu + torch.exp(0.5 * log_var) * torch.randn((S, mu.shape[1]), generator=g, device=dev, dtype=mu.dtype)
            recon += decode_batch_local(st, dec, z, L, g, mm).sum(dim=0, keepdim=True, dtype=torch.float64)
            done += S
        recon = (recon / num_samples).float()
    elif sampler == "stream":
        recon = 0
        for _ in range(num_samples):
            z = _draw(mu, log_var, g)
            recon = recon + decode_sample(st, dec, z, L, g, mm)
        recon = recon / num_samples
    else:
        raise ValueError(sampler)
    prior = torch.full((full_sequence_len, A + 5), -np.inf, dtype=torch.float32, device=dev)
    rows = torch.tensor([MSA_start + c for c in focus_cols], dtype=torch.long, device=dev)
    prior[rows, 5:] = recon[0]
    return prior


def cache_location(EVE_model_path: str, num_samples: int) -> str:
    parts = EVE_model_path.split("/")
    return "/".join(parts[:-1]) + os.sep + "log_prior" + os.sep + "_".join([parts[-1], str(num_samples), "log_space"])


def eve_log_prior(EVE_model_paths, EVE_model_parameters_location: str, msa, full_sequence_len: int, MSA_start: int,
                  EVE_num_samples_log_proba: int = 10, device="cuda", sampler: str = "auto") -> torch.Tensor:
    """get_EVE_models_and_log_prior (model_pytorch.py:940-967): ensemble mean of the per-model priors, each read from / written to
    the reference's cache location. ``msa`` is the MSAProcessing of the retrieved alignment (focus columns, trimmed focus sequence)."""
    from . import sharding
    params = json.load(open(EVE_model_parameters_location))
    rank, _ = sharding.rank_world()
    # Under torchrun every rank builds a scorer: rank 0 alone computes (and atomically writes) missing priors, the others wait at
    # the barrier and then read the cache — no rank ever unpickles a half-written file, and the Monte-Carlo pass runs once.
    if rank == 0:
        for path in EVE_model_paths:
            loc = cache_location(path, EVE_num_samples_log_proba)
            os.makedirs(os.path.dirname(loc), exist_ok=True)
            if not os.path.exists(loc):
                print("Computing EVE log prior")
                ck = torch.load(path, map_location="cpu")
                single = eve_log_prior_single(ck["model_state_dict"], params, msa.focus_seq_trimmed, msa.focus_cols, full_sequence_len,
                                              MSA_start, EVE_num_samples_log_proba, device, sampler=sampler).cpu()

                def dump(tmp, obj=single):
                    with open(tmp, "wb") as fh:
                        pickle.dump(obj, fh)
                sharding.atomic_write(loc, dump)
    sharding.barrier_if_distributed()
    total = 0
    for path in EVE_model_paths:
        loc = cache_location(path, EVE_num_samples_log_proba)
        print("Loading EVE log prior from disk")
        with open(loc, "rb") as fh:
            single = torch.as_tensor(pickle.load(fh))
        total = total + single.to(device)
    return total / len(EVE_model_paths)
