"""k-means and batched tokenisation (§8f #2): oracle vs the golden vectors of sklearn /
the reference's per-chunk loop on CPU, and the CUDA path vs both on the GPU box."""
import os

import numpy as np
import pytest
import torch

from make_kmeans_golden import KM_CASES, TOK, tok_inputs
from oracle import kmeans_oracle as KO
from oracle import vq_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _km_gold():
    return np.load(os.path.join(GOLD, "kmeans_lloyd.npz"))


def _tok_gold():
    g = np.load(os.path.join(GOLD, "tokenize_loop.npz"))
    hidden, E = tok_inputs()
    assert float(hidden.astype(np.float64).sum()) == float(g["hidden_sum"]), "RNG stream changed"
    assert float(E.astype(np.float64).sum()) == float(g["E_sum"]), "RNG stream changed"
    return g, hidden, E


# ---------------------------------------------------------------------------------------------
# CPU: the oracle is pinned to sklearn / the reference loop; host-side layout logic
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(KM_CASES))
def test_kmeans_oracle_matches_sklearn(name):
    c, g = KM_CASES[name], _km_gold()
    X, init = KO.synth_blobs(c["n"], c["d"], c["k"], c["seed"])
    assert float(X.astype(np.float64).sum()) == float(g[f"{name}_x_sum"]), "RNG stream changed"
    C, labels, inertia, n_iter = KO.lloyd(X, init, c["max_iter"], c["tol"])
    assert n_iter == int(g[f"{name}_n_iter"])
    assert np.array_equal(labels, g[f"{name}_labels"])
    np.testing.assert_allclose(C, g[f"{name}_centers"], rtol=1e-4, atol=1e-5)       # sklearn accumulates in fp32
    np.testing.assert_allclose(inertia, float(g[f"{name}_inertia"]), rtol=1e-5)
    Xh, _ = KO.synth_blobs(500, c["d"], c["k"], c["seed"] + 1000)
    assert np.array_equal(KO.assign(Xh, C), g[f"{name}_predict_heldout"])


def test_chunk_rows_layout_is_the_batch_of_one_view():
    from gesture2vec_b200.tokenizer import chunk_rows_from_hidden
    _, hidden, _ = _tok_gold()
    rows = chunk_rows_from_hidden(hidden)
    L, B, H = hidden.shape
    for b in (0, 1, B // 2, B - 1):          # what inputs.view(-1, D) gives for the [n_layers, 1, hidden] of chunk b
        assert np.array_equal(rows[b], np.ascontiguousarray(hidden[:, b:b + 1, :]).reshape(-1, L * H)[0])
    rt = chunk_rows_from_hidden(torch.from_numpy(hidden))
    assert np.array_equal(rt.numpy(), rows)
    # and it is NOT the flat view of the batched tensor (that pairs adjacent batch items)
    assert not np.array_equal(hidden.reshape(-1, L * H)[:B], rows)


def test_tokenize_oracle_matches_reference_loop():
    g, hidden, E = _tok_gold()
    from gesture2vec_b200.tokenizer import chunk_rows_from_hidden
    rows = chunk_rows_from_hidden(hidden)
    ids = O.nearest_code_f64(rows, E)
    aud = O.audit_indices(rows, E, ids, g["ids"])
    assert aud["hard"] == 0 and aud["mismatch"] == 0, aud


def test_kmeans_plusplus_seeding_is_sklearns():
    """The greedy k-means++ draw, row for row, against sklearn.cluster.kmeans_plusplus with the same RandomState
    (torch ops on CPU tensors here; the same code runs on the device in KMeans.fit).  The scikit-learn in this
    image is 1.3+ (first centre by `choice`); the reference's pinned 1.2.2 differs in that one draw (`randint`)."""
    from sklearn.cluster import kmeans_plusplus as sk_pp
    from gesture2vec_b200.kmeans import kmeans_plusplus
    rng = np.random.default_rng(0)
    for n, d, k in ((2000, 16, 12), (3000, 400, 20), (6000, 40, 64)):
        X = (rng.standard_normal((n, d)) + rng.integers(0, 5, (n, 1))).astype(np.float32)
        c_sk, idx_sk = sk_pp(X, k, random_state=np.random.RandomState(0))
        c, idx = kmeans_plusplus(torch.from_numpy(X), k, np.random.RandomState(0), seeding="sklearn-1.3+")
        assert np.array_equal(idx, idx_sk) and np.array_equal(c.numpy(), c_sk)
        # the 1.2 stream: same algorithm, first draw by randint -- a valid, reproducible seeding
        c2, idx2 = kmeans_plusplus(torch.from_numpy(X), k, np.random.RandomState(0), seeding="sklearn-1.2")
        c3, idx3 = kmeans_plusplus(torch.from_numpy(X), k, np.random.RandomState(0), seeding="sklearn-1.2")
        assert np.array_equal(idx2, idx3) and len(set(idx2.tolist())) == k and idx2[0] == np.random.RandomState(0).randint(n)


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_kmeans_fit_as_the_reference_calls_it_matches_sklearn():
    """Clustering.py:718 / train_DAE.py:257: KMeans(n_clusters, max_iter, random_state=0) -- k-means++ restarts from
    one RandomState stream, best inertia wins.  Same seeds as the scikit-learn in this image -> the same fit."""
    from sklearn.cluster import KMeans as SkKMeans
    import gesture2vec_b200 as g2v
    X, _ = KO.synth_blobs(6000, 40, 24, seed=11)
    sk = SkKMeans(n_clusters=24, n_init=4, max_iter=100, random_state=0, algorithm="lloyd").fit(X)
    km = g2v.KMeans(n_clusters=24, n_init=4, max_iter=100, random_state=0, seeding="sklearn-1.3+").fit(X)
    np.testing.assert_allclose(km.inertia_, sk.inertia_, rtol=1e-4)
    assert float((km.labels_ == sk.labels_).mean()) > 0.999
    np.testing.assert_allclose(km.cluster_centers_, sk.cluster_centers_, rtol=1e-3, atol=1e-3)
    # the lagged, sync-free convergence mode reaches the same fit
    km2 = g2v.KMeans(n_clusters=24, n_init=4, max_iter=100, random_state=0, seeding="sklearn-1.3+",
                     relocate_empty=False).fit(X)
    np.testing.assert_allclose(km2.inertia_, km.inertia_, rtol=1e-5)
    assert km2.n_iter_ == km.n_iter_ and np.array_equal(km2.labels_, km.labels_)


@pytest.mark.gpu
def test_kmeans_relocates_empty_clusters():
    """A centre no row is assigned to is moved onto the row farthest from its centre (sklearn's
    _relocate_empty_clusters_dense), instead of staying empty for ever."""
    from sklearn.cluster import KMeans as SkKMeans
    import gesture2vec_b200 as g2v
    X, init = KO.synth_blobs(3000, 16, 8, seed=3)
    init = init.copy()
    init[5] = 1e3                                   # far from every row: empty after the first assignment
    km = g2v.KMeans(n_clusters=8, init=init, max_iter=50).fit(X)
    assert np.bincount(km.labels_, minlength=8).min() > 0
    sk = SkKMeans(n_clusters=8, init=init, n_init=1, max_iter=50, algorithm="lloyd").fit(X)
    np.testing.assert_allclose(km.inertia_, sk.inertia_, rtol=0.05)     # which far row is taken may differ
    stay = g2v.KMeans(n_clusters=8, init=init, max_iter=50, relocate_empty=False).fit(X)
    assert np.bincount(stay.labels_, minlength=8).min() == 0 and stay.inertia_ > km.inertia_


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(KM_CASES))
def test_kmeans_fit_matches_sklearn_golden(name):
    import gesture2vec_b200 as g2v
    c, g = KM_CASES[name], _km_gold()
    X, init = KO.synth_blobs(c["n"], c["d"], c["k"], c["seed"])
    km = g2v.KMeans(n_clusters=c["k"], init=init, max_iter=c["max_iter"], tol=c["tol"]).fit(X)
    assert km.n_iter_ == int(g[f"{name}_n_iter"])
    assert km.labels_.dtype == np.int32 and km.cluster_centers_.shape == (c["k"], c["d"])
    assert np.array_equal(km.labels_, g[f"{name}_labels"])
    # centres are per-cluster means: fp32 atomics (ours) vs fp32 chunked sums (sklearn) -> 1e-4
    np.testing.assert_allclose(km.cluster_centers_, g[f"{name}_centers"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(km.inertia_, float(g[f"{name}_inertia"]), rtol=1e-5)
    Xh, _ = KO.synth_blobs(500, c["d"], c["k"], c["seed"] + 1000)
    assert np.array_equal(km.predict(Xh), g[f"{name}_predict_heldout"])
    # a model fitted elsewhere (e.g. an unpickled sklearn KMeans) only needs its centres
    km2 = g2v.KMeans.from_centers(g[f"{name}_centers"])
    assert np.array_equal(km2.predict(torch.from_numpy(Xh).cuda()), g[f"{name}_predict_heldout"])


@pytest.mark.gpu
def test_kmeans_default_seeding_and_lloyd_properties():
    import gesture2vec_b200 as g2v
    X, _ = KO.synth_blobs(20000, 400, 300, seed=5)          # Clustering.py:718: 300 clusters on [N, 400] latents
    one = g2v.KMeans(n_clusters=300, max_iter=1, random_state=0).fit(X)
    km = g2v.KMeans(n_clusters=300, max_iter=40, random_state=0).fit(X)
    assert km.inertia_ <= one.inertia_ * (1 + 1e-6)          # Lloyd never increases the inertia
    assert km.n_iter_ <= 40 and km.labels_.shape == (20000,)
    # fixed point: the labels are the exact assignment to the final centres, the centres the means of their rows
    ref = KO.assign(X, km.cluster_centers_)
    aud = O.audit_indices(X, km.cluster_centers_, km.labels_.astype(np.int64), ref.astype(np.int64))
    assert aud["hard"] == 0, aud
    again = g2v.KMeans(n_clusters=300, max_iter=40, random_state=0).fit(X)      # same seed, same result
    assert np.array_equal(again.labels_, km.labels_)
    with pytest.raises(ValueError):
        g2v.KMeans(n_clusters=8).fit(X[:4])


@pytest.mark.gpu
def test_tokenizer_matches_reference_loop_golden():
    import gesture2vec_b200 as g2v
    g, hidden, E = _tok_gold()
    layer = g2v.DAE_VQ_Payam(TOK["K"], TOK["n_layers"] * TOK["hidden"], 0.25)
    with torch.no_grad():
        layer._embedding.weight.copy_(torch.from_numpy(E))
    tok = g2v.GestureTokenizer(layer.cuda())
    ids = tok.encode_hidden(hidden)                                   # host array -> pinned-buffer entry point
    assert ids.dtype == np.int64 and np.array_equal(ids, g["ids"])
    ids_dev = tok.encode_hidden(torch.from_numpy(hidden).cuda())      # device tensor -> searched in place
    assert np.array_equal(ids_dev, g["ids"])
    entries = tok.clustering_entries([hidden[:, b:b + 1, :] for b in range(hidden.shape[1])])
    assert len(entries) == hidden.shape[1]
    assert entries[3]["quantized_indices"].shape == (1,) and entries[3]["quantized_indices"][0] == g["ids"][3]
    assert entries[3]["latent_rnn"].shape == (TOK["n_layers"], TOK["hidden"])
    cid = tok.cluster_ids(g2v.chunk_rows_from_hidden(hidden))
    assert cid.dtype == torch.int64 and np.array_equal(cid.numpy(), g["ids"])
    # the same ids as the drop-in module called chunk by chunk, as Clustering.py does
    for b in (0, 100, 256):
        _, _, _, enc = layer(torch.from_numpy(np.ascontiguousarray(hidden[:, b:b + 1, :])).cuda())
        assert int(torch.argmax(enc, 1)) == int(g["ids"][b])
    assert tok.encode_rows(np.zeros((0, tok.D), np.float32)).shape == (0,)
