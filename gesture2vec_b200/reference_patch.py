"""Make the reference's scripts use the B200 quantizers without editing the reference.

The reference builds its quantizer inside ``Autoencoder_VQVAE.__init__``
(scripts/model/Autoencoder_VQVAE_model.py:794-820: VQ_Payam_EMA, then unconditionally
overwritten by the soft VQ_Payam_GSSoft) and inside ``VQ_Frame.__init__``
(scripts/model/DAE_model.py:162-176).  Two hooks cover both:

  patch_reference()        rebinds model.Autoencoder_VQVAE_model.{VQ_Payam,VQ_Payam_EMA,
                           VectorQuantizerEMA,VQ_Payam_GSSoft} and model.DAE_model.{VQ_Payam,VQ_Payam_EMA} to the
                           classes of gesture2vec_b200.quantizers, so every later construction
                           (train_autoencoder_VQVAE.init_model, utils/train_utils.load_checkpoint_and_model)
                           gets the CUDA-backed layer with identical state_dict keys.
  swap_vq_layer(net, ...)  replaces an already-built ``net.vq_layer`` (whatever class it is) by a
                           hard quantizer, copying the codebook / EMA state / pre_linear when the
                           shapes agree -- needed because the shipped __init__ ends on GSSoft.
"""
from __future__ import annotations

import importlib
from typing import Optional

import torch

from . import quantizers as Q

_TARGETS = {
    "model.Autoencoder_VQVAE_model": Q.FLAVOURS["vqvae"],
    "model.DAE_model": Q.FLAVOURS["dae"],
}
_saved = {}


def patch_reference(modules: Optional[dict] = None) -> dict:
    """Rebind the reference's hard quantizer classes.  `modules` may map module name -> module
    object (already imported); otherwise they are imported (scripts/ must be on sys.path).
    Returns {module_name: [patched class names]}."""
    done = {}
    for modname, classes in _TARGETS.items():
        mod = (modules or {}).get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
        for cname, cls in classes.items():
            if hasattr(mod, cname):
                _saved.setdefault((modname, cname), getattr(mod, cname))
                setattr(mod, cname, cls)
                done.setdefault(modname, []).append(cname)
    return done


def unpatch_reference(modules: Optional[dict] = None) -> None:
    for (modname, cname), orig in list(_saved.items()):
        mod = (modules or {}).get(modname) or importlib.import_module(modname)
        setattr(mod, cname, orig)
        del _saved[(modname, cname)]


@torch.no_grad()
def swap_vq_layer(net: torch.nn.Module, kind: str = "VQ_Payam_EMA", flavour: str = "vqvae",
                  decay: float = 0.85, attr: str = "vq_layer") -> torch.nn.Module:
    """Replace ``net.<attr>`` by a CUDA-backed hard quantizer of the given kind, keeping state."""
    old = getattr(net, attr)
    K, D = old._num_embeddings, old._embedding_dim
    beta = float(old._commitment_cost)
    cls = Q.FLAVOURS[flavour][kind]
    new = cls(K, D, beta, getattr(old, "_decay", decay), getattr(old, "_epsilon", 1e-5)) \
        if kind not in ("VQ_Payam", "VQ_Payam_GSSoft") else cls(K, D, beta)
    src = old.state_dict()
    dst = new.state_dict()
    for k, v in src.items():
        if k in dst and dst[k].shape == v.shape:
            dst[k].copy_(v)
    new.load_state_dict(dst)
    new.to(old._embedding.weight.device)
    new.train(old.training)
    setattr(net, attr, new)
    return new
