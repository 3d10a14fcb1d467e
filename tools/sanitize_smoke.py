"""One small invocation of every kernel variant, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitize_smoke.py

Sizes are tiny (the tools slow kernels down 10-100x) but cover: both tcgen05 sweeps in all their variants (CTA
pairs and single CTA, fp32 / bf16 / fp16 rows, ragged N, K and D tails), the fp32 sweep, the exact re-rank lists,
the row pass with and without residual sums, the cooperative finalise launch in every mode, backward, one-hot,
the fp64 checker, the host-buffer tokeniser and one Lloyd iteration.  Results are checked against the fp64 checker.
"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import gesture2vec_b200 as g  # noqa: E402
from gesture2vec_b200 import _lib as L  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    variants = {"auto": L.ALGO_AUTO, "simt": L.ALGO_SIMT, "tmem": L.ALGO_TC | L.TC_VARIANT_TMEM,
                "fused": L.ALGO_TC | L.TC_VARIANT_FUSED, "prep": L.ALGO_TC | L.TC_VARIANT_PREP}
    n_checked = 0
    for (N, K, D) in ((600, 400, 400), (257, 512, 400), (300, 1000, 104), (130, 80, 40)):
        E = torch.randn(K, D, device=dev, generator=gen)
        z = torch.randn(N, D, device=dev, generator=gen)
        exact = g.vq_search_exact(z, E)
        cb = g.prepare_codebook(E)
        for name, flags in variants.items():
            if name != "simt" and L.load().g2v_search_path(K, D, 0) == L.ALGO_SIMT:
                continue
            for zz in (z, z.to(torch.bfloat16), z.to(torch.float16)):
                ex = exact if zz.dtype == torch.float32 else g.vq_search_exact(zz, E)
                idx = g.vq_search(zz, E, cb, flags=flags)
                bad = int((idx != ex).sum())
                assert bad <= 2, (N, K, D, name, zz.dtype, bad)      # exact-arithmetic ties aside
                n_checked += 1
    # training step of both EMA flavours + hard VQ, eval step, tokenisers, k-means iteration
    for cls in (g.DAE_VQ_Payam_EMA, g.VQVAE_VQ_Payam_EMA):
        layer = cls(512, 400, 0.25, 0.85).to(dev).train()
        x = torch.tanh(torch.randn(2, 128, 200, device=dev, generator=gen)).requires_grad_(True)
        for inplace in (False, True):
            layer.ema_inplace = inplace
            loss, q, ppl, enc = layer(x)
            (loss + q.sum()).backward()
        layer.eval()
        with torch.no_grad():
            layer(x)
        layer.tokenize(x.detach())
    hard = g.DAE_VQ_Payam(80, 40, 0.25).to(dev)
    x = torch.randn(256, 40, device=dev, generator=gen, requires_grad=True)
    loss, q, ppl, enc = hard(x)
    (loss + q.sum()).backward()
    zh = torch.randn(5000, 400, generator=torch.Generator().manual_seed(1)).pin_memory()
    E = torch.randn(400, 400, device=dev, generator=gen)
    ids = g.tokenize_host(zh, E, chunk_rows=2048)
    assert int((ids.to(dev) != g.vq_search_exact(zh.to(dev), E)).sum()) <= 2
    g.KMeans(n_clusters=16, init=zh[:16].numpy(), max_iter=2, device=dev).fit(zh[:3000].numpy())
    # the refine pass of whole-row re-ranks (side stream: split operands, tcgen05 GEMM with a device-side row count,
    # candidate selection) and the flat backward / cp.async row pass at a bulk size: every code four times, so every
    # row ties in two of the 32 column chains and is listed as a whole row
    E = torch.randn(48, 128, device=dev, generator=gen).repeat(4, 1)
    z = torch.randn(32768, 128, device=dev, generator=gen)
    st = torch.zeros(8, dtype=torch.int64, device=dev)
    idx = g.vq_search(z, E, stats=st)
    assert int(st[L.STAT_REFINE_ROWS]) > 0, st.tolist()
    assert int((idx != g.vq_search_exact(z, E)).sum()) == 0
    layer = g.DAE_VQ_Payam_EMA(256, 64, 0.25, 0.85).to(dev).train()
    xb = torch.randn(32768, 64, device=dev, generator=gen, requires_grad=True)
    loss, q, ppl, enc = layer(xb)
    (loss + q.sum()).backward()
    soft = g.VQVAE_VQ_Payam_GSSoft(64, 64, 0.25).to(dev)
    xs = torch.randn(4, 64, 32, device=dev, generator=gen, requires_grad=True)
    loss, q, ppl, enc = soft(xs)
    (loss + q.sum()).backward()
    torch.cuda.synchronize()
    print(f"sanitize smoke ok: {n_checked} search variants checked")


if __name__ == "__main__":
    main()
