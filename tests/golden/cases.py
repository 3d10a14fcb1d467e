"""Case table + seeded input regeneration shared by make_golden.py and the tests.

Inputs are NOT stored in the fixtures: they are regenerated from numpy PCG64 seeds
(`regen`).  Every fixture stores an fp64 checksum of each regenerated array and
`load_case` asserts it, so an RNG-stream change fails loudly instead of silently.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle.vq_oracle import synth_codebook, synth_latents  # noqa: E402

# name, flavour, class, input shape, K, D, latent kind, codebook kind, decay
CASES = [
    ("dae_hard_frame40",    "dae",   "VQ_Payam",     (256, 40),     80,  40, "gru",       "uniform_invK", None),
    ("dae_ema_frame45",     "dae",   "VQ_Payam_EMA", (250, 45),     64,  45, "iid",       "uniform1",     0.99),
    ("vqvae_hard_trinity",  "vqvae", "VQ_Payam",     (2, 128, 200), 512, 400, "iid",      "normal",       None),
    ("vqvae_ema_trinity",   "vqvae", "VQ_Payam_EMA", (2, 128, 200), 512, 400, "gru",      "uniform1",     0.85),
    ("dae_ema_genea",       "dae",   "VQ_Payam_EMA", (2, 64, 200),  400, 400, "clustered", "normal",      0.85),
    ("dae_hard_batch1",     "dae",   "VQ_Payam",     (2, 1, 200),   400, 400, "gru",      "uniform1",     None),
    ("dae_hard_dupcodes",   "dae",   "VQ_Payam",     (64, 40),      32,  40, "iid",       "normal",       None),
]
BETA = 0.25
EPS = 1e-5
EMA_STEPS = 3
G_LOSS = 3.0          # the fixtures back-propagate  G_LOSS*loss + sum(quantized*g_out)
CASE_NAMES = [c[0] for c in CASES]


def case_spec(name):
    ci = CASE_NAMES.index(name)
    return ci, CASES[ci]


def regen(name):
    """All inputs of a case, deterministically from seeds."""
    ci, (_, flavour, cls, shape, K, D, lkind, ckind, decay) = case_spec(name)
    ema = cls.endswith("EMA")
    E0 = synth_codebook(ckind, K, D, seed=100 + ci)
    if name.endswith("dupcodes"):            # exact ties: duplicate codebook rows
        E0[1::2] = E0[0::2]
    rng = np.random.default_rng(500 + ci)
    r = dict(E0=E0, K=K, D=D, shape=shape, flavour=flavour, cls=cls, ema=ema, decay=decay,
             beta=BETA, eps=EPS, steps=EMA_STEPS if ema else 1)
    if ema:
        r["ema_w0"] = rng.standard_normal((K, D), dtype=np.float32)
    if flavour == "vqvae" or ema:            # classes that own a pre_linear
        bound = 1.0 / np.sqrt(D)
        r["pre_W"] = rng.uniform(-bound, bound, (D, D)).astype(np.float32)
        r["pre_b"] = rng.uniform(-bound, bound, (D,)).astype(np.float32)
    n = int(np.prod(shape)) // D
    for s in range(r["steps"] + (1 if ema else 0)):   # last one = eval-mode input for EMA cases
        r[f"x{s}"] = synth_latents(lkind, n, D, E=E0, seed=900 + 10 * ci + s).reshape(shape)
        r[f"g{s}"] = rng.standard_normal(shape, dtype=np.float32)
    return r


def checksums(r):
    return {k: float(np.asarray(v, np.float64).sum()) for k, v in r.items()
            if isinstance(v, np.ndarray)}


def load_case(name):
    """(inputs, golden) with the regeneration checksums verified."""
    g = np.load(os.path.join(HERE, name + ".npz"), allow_pickle=False)
    r = regen(name)
    for k, v in checksums(r).items():
        ref = float(g["chk_" + k])
        assert abs(v - ref) <= 1e-9 * max(1.0, abs(ref)), f"regenerated input {k} drifted"
    return r, g
