"""CPU integration test against the REAL reference classes (skips where /root/reference is absent, e.g. on the GPU
box): build Autoencoder_VQVAE from a hand-made Namespace the way train_autoencoder_VQVAE.init_model does, with and
without patch_reference(), swap the quantizer, and strict-load state dicts in both directions."""
import argparse
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/scripts"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not on this machine")


@pytest.fixture(scope="module")
def ref():
    shim = types.ModuleType("configargparse")
    shim.argparse = argparse
    sys.modules.setdefault("configargparse", shim)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model.Autoencoder_VQVAE_model as vqvae
    import model.DAE_model as dae
    vqvae.debug = False
    return vqvae, dae


def _args(**kw):
    a = dict(rep_learning_dim=40, hidden_size=200, n_layers=2, dropout_prob=0.1, autoencoder_vae="False",
             autoencoder_vq="True", autoencoder_vq_components=512, autoencoder_vq_commitment_cost=0.25, n_pre_poses=4,
             autoencoder_conditioned="True", autoencoder_att="False", autoencoder_fixed_weight="False", n_poses=34,
             motion_resampling_framerate=20, wordembed_dim=300, z_type="none", input_context="none")
    a.update(kw)
    return argparse.Namespace(**a)


def test_real_autoencoder_constructs_with_the_drop_in_classes(ref):
    import gesture2vec_b200 as g
    vqvae, dae = ref
    stock = vqvae.Autoencoder_VQVAE(_args(), pose_dim=135, n_frames=30)
    assert type(stock.vq_layer).__name__ == "VQ_Payam_GSSoft"            # the shipped __init__ ends on the soft layer
    done = g.patch_reference()
    try:
        assert set(done["model.Autoencoder_VQVAE_model"]) >= {"VQ_Payam", "VQ_Payam_EMA", "VQ_Payam_GSSoft"}
        assert set(done["model.DAE_model"]) == {"VQ_Payam", "VQ_Payam_EMA"}
        net = vqvae.Autoencoder_VQVAE(_args(), pose_dim=135, n_frames=30)
        assert isinstance(net.vq_layer, g.VQVAE_VQ_Payam_GSSoft)
        # identical state_dict contract: the stock checkpoint strict-loads into the patched model and back
        net.load_state_dict(stock.state_dict(), strict=True)
        stock.load_state_dict(net.state_dict(), strict=True)
        assert torch.equal(net.vq_layer._embedding.weight, stock.vq_layer._embedding.weight)
    finally:
        g.unpatch_reference()
    assert vqvae.VQ_Payam_EMA is not g.VQVAE_VQ_Payam_EMA


def test_swap_keeps_state_and_hard_quantizers_strict_load(ref):
    import gesture2vec_b200 as g
    vqvae, dae = ref
    net = vqvae.Autoencoder_VQVAE(_args(autoencoder_vq_components=400), pose_dim=135, n_frames=30)
    soft_E = net.vq_layer._embedding.weight.detach().clone()
    new = g.swap_vq_layer(net, kind="VQ_Payam_EMA", flavour="vqvae", decay=0.85)
    assert net.vq_layer is new and isinstance(new, g.VQVAE_VQ_Payam_EMA)
    assert torch.equal(new._embedding.weight.detach(), soft_E) and new._num_embeddings == 400 and new._embedding_dim == 400
    for ref_cls, ours in ((vqvae.VQ_Payam_EMA, g.VQVAE_VQ_Payam_EMA), (dae.VQ_Payam_EMA, g.DAE_VQ_Payam_EMA)):
        r, o = ref_cls(64, 40, 0.25, 0.9), ours(64, 40, 0.25, 0.9)
        o.load_state_dict(r.state_dict(), strict=True)
        r.load_state_dict(o.state_dict(), strict=True)
        assert o._decay == r._decay and o._epsilon == r._epsilon and o._commitment_cost == r._commitment_cost
    for ref_cls, ours in ((vqvae.VQ_Payam, g.VQVAE_VQ_Payam), (dae.VQ_Payam, g.DAE_VQ_Payam),
                          (vqvae.VQ_Payam_GSSoft, g.VQVAE_VQ_Payam_GSSoft)):
        r, o = ref_cls(64, 40, 0.25), ours(64, 40, 0.25)
        o.load_state_dict(r.state_dict(), strict=True)
        r.load_state_dict(o.state_dict(), strict=True)
    # DAE: VQ_Frame builds its quantizer through the module-level names too
    g.patch_reference()
    try:
        assert dae.VQ_Payam_EMA is g.DAE_VQ_Payam_EMA
    finally:
        g.unpatch_reference()
