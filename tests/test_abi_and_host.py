"""CPU-only checks: the C-ABI library loads and exports every symbol include/g2v_vq.h declares,
argument validation that needs no GPU, and the host-side logic (sharding, packed layout,
reference patching, module construction / state_dict contract)."""
import ctypes
import json
import os
import re
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib():
    from gesture2vec_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from gesture2vec_b200.build import build
        build()
    return _lib.load()


def test_header_symbols_all_exported_and_bound(lib):
    from gesture2vec_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "g2v_vq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(g2v_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed from the header"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} not exported by the shared library"


def test_version_and_error_strings(lib):
    assert lib.g2v_version() == 200
    assert lib.g2v_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert lib.g2v_strerror(code) not in (b"ok", b"unknown error")
    assert lib.g2v_strerror(-99) == b"unknown error"


def test_size_queries_and_validation_without_gpu(lib):
    assert lib.g2v_codebook_bytes(0, 4) == 0
    n = lib.g2v_codebook_bytes(400, 400)
    assert n >= 400 * 4 + 512 * 400 * 2 and n % 256 == 0
    assert lib.g2v_workspace_bytes(-1, 4, 4, 0, 0) == 0
    assert lib.g2v_workspace_bytes(1000, 512, 400, 0, 1) >= 1000 * 4          # SIMT: the re-rank list
    assert lib.g2v_workspace_bytes(1000, 512, 400, 0, 0) >= 1000 * 400 * 2      # TC: fp16 operand rows
    # bulk searches reserve the refine pass's region (split operands + fp32 dots of the listed rows), capped at 4 GiB;
    # batches below 32768 rows do not
    small, bulk = lib.g2v_workspace_bytes(32767, 512, 400, 0, 0), lib.g2v_workspace_bytes(32768, 512, 400, 0, 0)
    assert bulk - small > 32768 * 512 * 4
    assert lib.g2v_workspace_bytes(1 << 24, 512, 400, 0, 0) - lib.g2v_workspace_bytes(1 << 23, 512, 400, 0, 0) < (1 << 23) * 2000
    # the sorted row pass: scratch only for bulk batches and K <= 4096
    assert lib.g2v_apply_workspace_bytes(16383, 512) == 0 and lib.g2v_apply_workspace_bytes(1 << 20, 8192) == 0
    assert lib.g2v_apply_workspace_bytes(1 << 20, 512) >= (1 << 20) * 4 + 2 * 512 * 4
    assert lib.g2v_vq_apply_ws(None, None, None, None, 10, 4, 4, None, None, None, None, 0, None, 0, None) == -1
    assert lib.g2v_search_path(512, 400, 0) == 2 and lib.g2v_search_path(512, 400, 1) == 1
    assert lib.g2v_search_path(80, 40, 0) == 1                                  # tiny frame-level shape -> fp32 path
    assert lib.g2v_search_path(80, 40, 2) == -7
    assert lib.g2v_search_path(0, 40, 0) == -1
    # null / bad arguments are rejected before any CUDA call
    assert lib.g2v_vq_search(None, 0, None, None, 10, 4, 4, None, None, None, 0, 0, None) == -1
    assert lib.g2v_vq_apply(None, None, None, None, 10, 4, 4, None, None, None, None, 0, None) == -1
    assert lib.g2v_onehot(None, 5, 0, None, None) == -1
    assert lib.g2v_vq_stats_pack(None, None, None, 0, 1, 4, 4, None, None) == -1
    assert lib.g2v_tokenize_host_bytes(0, 4, 4, 0, 0) == 0
    # the fused finalise: bad update mode, EMA with aliased cluster sizes, missing statistics buffer
    assert lib.g2v_vq_step_finalize(None, None, None, 0, 0, None, 4, 4, 0.0, 0.0, None, None, 7, None, None, None, None,
                                    None, None, None, 0.0, 0.0, None, None, 0, None) == -1
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.g2v_vq_step_finalize(None, None, None, 0, 0, p, 4, 4, 0.0, 0.0, None, None, 1, p, p, p, p, p, p, p, 0.9,
                                    1e-5, None, None, 0, None) == -1    # E_prev aliases the codebook
    assert lib.g2v_vq_ema_update(p, p, p, p, None, p, p, 0.9, 1e-5, 4, 4, None, 0, None) == -1
    assert lib.g2v_exact_workspace_bytes(400) >= 400 * 8 and lib.g2v_exact_workspace_bytes(0) == 0
    assert lib.g2v_vq_search_exact(None, 0, None, 10, 4, 4, None, None, 0, None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from gesture2vec_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libg2v_vq.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_shard_rows_partition():
    from gesture2vec_b200 import shard_rows
    for n in (0, 1, 7, 1_000_000, 16_777_216):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_packed_layout_matches_header_contract():
    from gesture2vec_b200 import packed_layout, packed_numel
    lay = packed_layout(512, 400)
    assert lay["dwr"] == (0, 204800) and lay["counts"] == (204800, 205312)
    assert lay["sse"] == 205312 and lay["rows"] == 205313 and lay["numel"] == packed_numel(512, 400) == 205314
    assert lay["numel"] * 4 == 821256          # bytes all-reduced per EMA step at K=512


def test_modules_construct_with_reference_state_dict_keys():
    import gesture2vec_b200 as g
    keys = json.load(open(os.path.join(HERE, "golden", "state_dict_keys.json")))
    for flavour, classes in g.FLAVOURS.items():
        for cname, cls in classes.items():
            layer = cls(8, 4, 0.25) if cname in ("VQ_Payam", "VQ_Payam_GSSoft") else cls(8, 4, 0.25, 0.9)
            assert sorted(layer.state_dict().keys()) == keys[f"{flavour}.{cname}"]
            assert layer._num_embeddings == 8 and layer._embedding_dim == 4 and layer._commitment_cost == 0.25
            if cname not in ("VQ_Payam", "VQ_Payam_GSSoft"):
                assert layer._decay == 0.9 and layer._epsilon == 1e-5
                assert tuple(layer._ema_w.shape) == (8, 4) and tuple(layer._ema_cluster_size.shape) == (8,)
            layer.embedding_grad(False)
            assert not layer._embedding.weight.requires_grad


def test_cpu_module_raises_instead_of_falling_back():
    import gesture2vec_b200 as g
    layer = g.DAE_VQ_Payam(8, 4, 0.25)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.randn(3, 4))


def test_patch_and_swap_on_stand_in_modules():
    import gesture2vec_b200 as g
    fake_vq = types.ModuleType("model.Autoencoder_VQVAE_model")
    fake_dae = types.ModuleType("model.DAE_model")
    for m in (fake_vq, fake_dae):
        m.VQ_Payam = object
        m.VQ_Payam_EMA = object
    fake_vq.VectorQuantizerEMA = object
    done = g.patch_reference({"model.Autoencoder_VQVAE_model": fake_vq, "model.DAE_model": fake_dae})
    assert fake_vq.VQ_Payam_EMA is g.VQVAE_VQ_Payam_EMA and fake_dae.VQ_Payam is g.DAE_VQ_Payam
    assert sorted(done["model.Autoencoder_VQVAE_model"]) == ["VQ_Payam", "VQ_Payam_EMA", "VectorQuantizerEMA"]
    g.unpatch_reference({"model.Autoencoder_VQVAE_model": fake_vq, "model.DAE_model": fake_dae})
    assert fake_vq.VQ_Payam is object and fake_dae.VQ_Payam_EMA is object

    # swap an already-built layer (the shipped Autoencoder_VQVAE ends on the soft quantizer)
    class Soft(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._num_embeddings, self._embedding_dim, self._commitment_cost = 16, 8, 0.25
            self._embedding = torch.nn.Embedding(16, 8)
            self.pre_linear = torch.nn.Linear(8, 8)
    net = torch.nn.Module()
    net.vq_layer = Soft()
    old_E = net.vq_layer._embedding.weight.detach().clone()
    new = g.swap_vq_layer(net, kind="VQ_Payam_EMA", flavour="vqvae", decay=0.85)
    assert net.vq_layer is new and isinstance(new, g.VQVAE_VQ_Payam_EMA) and new._decay == 0.85
    assert torch.equal(new._embedding.weight.detach(), old_E)
