"""Device-side synthetic latents / codebooks for the large GPU audits (SURVEY.md 8d distributions), and the
near-tie classifier that compares two index vectors in fp64 on the device.  Test infrastructure only."""
import numpy as np
import torch

from oracle import vq_oracle as O


def codebook(kind: str, K: int, D: int, dev, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device=dev).manual_seed(seed)
    if kind == "normal":
        return torch.randn(K, D, device=dev, generator=g)
    if kind == "uniform1":
        return torch.rand(K, D, device=dev, generator=g) * 2 - 1
    if kind == "uniform_invK":
        return (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
    if kind == "ema_degenerate":
        # the codebook after ONE reference EMA step from the class init (unused codes blow up to |E| ~ 1e5)
        E0 = O.synth_codebook("uniform1", K, D, seed=seed)
        w0 = np.random.default_rng(seed + 1).standard_normal((K, D), dtype=np.float32)
        layer = O.EmaVQ(E0, w0, 0.25, 0.85)
        layer.forward(O.synth_latents("gru", 128, D, seed=seed + 2))
        return torch.from_numpy(np.ascontiguousarray(layer.E)).to(dev)
    raise ValueError(kind)


def latents(kind: str, N: int, D: int, dev, E: torch.Tensor = None, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator(device=dev).manual_seed(seed)
    if kind == "iid":
        return torch.randn(N, D, device=dev, generator=g)
    if kind == "gru":
        return torch.tanh(0.8 * torch.randn(N, D, device=dev, generator=g))
    if kind == "clustered":
        K = E.shape[0]
        w = 1.0 / torch.arange(1, K + 1, device=dev, dtype=torch.float64) ** 1.1
        c = torch.multinomial(w / w.sum(), N, replacement=True, generator=g)
        z = E[c]
        z = torch.where(torch.isfinite(z) & (z.abs() < 1e3), z, torch.zeros_like(z))
        return (z + 0.1 * torch.randn(N, D, device=dev, generator=g)).contiguous()
    raise ValueError(kind)


def audit(z: torch.Tensor, E: torch.Tensor, got: torch.Tensor, want: torch.Tensor, eps_tie: float = O.EPS_TIE) -> dict:
    """Rows where `got` != `want`, split into near-ties (fp64 gap of the two chosen codes <= eps_tie * (|z|^2 +
    |e|^2)) and hard mismatches; `wrong_side` counts mismatches where `got` is the strictly farther code by more
    than the tolerance (the hard ones) -- everything evaluated in fp64 on the device for the mismatching rows only."""
    bad = torch.nonzero(got.long() != want.long()).flatten()
    out = {"rows": int(z.shape[0]), "mismatch": int(bad.numel()), "near_tie": 0, "hard": 0, "max_gap_rel": 0.0}
    if bad.numel() == 0:
        return out
    zz = z[bad].double()
    ea, eb = E[got[bad].long()].double(), E[want[bad].long()].double()
    da, db = ((zz - ea) ** 2).sum(1), ((zz - eb) ** 2).sum(1)
    scale = (zz ** 2).sum(1) + torch.maximum((ea ** 2).sum(1), (eb ** 2).sum(1))
    rel = (da - db).abs() / scale
    hard = rel > eps_tie
    # an equal distance with a larger index is also a defect (first index must win exact ties)
    tie_order = (da == db) & (got[bad].long() > want[bad].long()) & (eps_tie < 2.0 ** -30)
    out["hard"] = int((hard | tie_order).sum())
    out["near_tie"] = out["mismatch"] - out["hard"]
    out["max_gap_rel"] = float(rel.max())
    return out
