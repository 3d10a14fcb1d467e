"""Full-row exact audit of every search kernel variant on the GPU.

    python tools/audit_exact.py [--rows 1000000] [--out gpurun_out/audit.log]

For each latent / codebook distribution of SURVEY.md 8d (iid, GRU-like, clustered, degenerate-EMA codebook):
ONE pass of the fp64 device checker (g2v_vq_search_exact) over all rows, then every kernel variant
(tc_tmem<f32>, tc_search<1,fused>, row_prep + tc_search<2>, fp32 SIMT, and the 16-bit-row variants on their own
exact answer) is compared with it row by row; mismatches are classified in fp64 (near-tie: gap below
2^-40 (|z|^2+|e|^2), i.e. exact-arithmetic ties only).  The log goes to stdout and --out.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gpu_synth as S  # noqa: E402
import gesture2vec_b200 as g  # noqa: E402
from gesture2vec_b200 import _lib as L  # noqa: E402

CASES = [("iid", "normal", 400), ("gru", "uniform1", 512), ("clustered", "normal", 400), ("gru", "ema_degenerate", 512),
         ("iid", "normal", 16384)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--rows-large-k", type=int, default=65536)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "audit_exact.log"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    D = 400
    variants = {"auto": L.ALGO_AUTO, "simt": L.ALGO_SIMT, "tmem": L.ALGO_TC | L.TC_VARIANT_TMEM,
                "fused": L.ALGO_TC | L.TC_VARIANT_FUSED, "prep": L.ALGO_TC | L.TC_VARIANT_PREP}
    lines = []

    def log(rec):
        s = json.dumps(rec)
        print(s, flush=True)
        lines.append(s)

    log({"what": "full-row fp64 audit", "device": torch.cuda.get_device_name(0), "rows": a.rows, "D": D,
         "eps_tie": "2^-40 (exact-arithmetic ties only)", "lib_version": L.load().g2v_version()})
    total_hard = 0
    for lk, ck, K in CASES:
        N = a.rows if K <= 2048 else a.rows_large_k
        E = S.codebook(ck, K, D, dev, seed=7)
        z = S.latents(lk, N, D, dev, E=E, seed=8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        exact = g.vq_search_exact(z, E)
        torch.cuda.synchronize()
        t_exact = time.perf_counter() - t0
        cb = g.prepare_codebook(E)
        names = list(variants) if K <= 2048 else ["auto", "prep", "simt"]
        for name in names:
            if name == "tmem" and K > 576:
                continue
            stats = torch.zeros(8, dtype=torch.int64, device=dev)
            idx = g.vq_search(z, E, cb, flags=variants[name], stats=stats)
            r = S.audit(z, E, idx, exact, eps_tie=2.0 ** -40)
            st = stats.cpu().tolist()
            r.update(latents=lk, codebook=ck, K=K, dtype="f32", variant=name, exact_checker_s=round(t_exact, 2),
                     rerank_rows=st[1], fallback_rows=st[3], fp64_rows=st[2])
            total_hard += r["hard"]
            log(r)
        # every row through the refine pass (split-fp16 dots, a-priori bound, fp64 on what the bound cannot exclude)
        if K <= 2048:
            stats = torch.zeros(8, dtype=torch.int64, device=dev)
            idx = g.vq_search(z, E, cb, flags=L.ALGO_TC | L.LIST_ALL_ROWS, stats=stats)
            r = S.audit(z, E, idx, exact, eps_tie=2.0 ** -40)
            st = stats.cpu().tolist()
            r.update(latents=lk, codebook=ck, K=K, dtype="f32", variant="refine_all_rows", refine_rows=st[4],
                     refine_fp64_codes=st[5])
            total_hard += r["hard"]
            log(r)
        n16 = min(N, 262144)
        for dt in (torch.bfloat16, torch.float16):
            z16 = z[:n16].to(dt).contiguous()
            ex16 = g.vq_search_exact(z16, E)
            for name in ("auto", "prep", "simt"):
                idx = g.vq_search(z16, E, cb, flags=variants[name])
                r = S.audit(z16.float(), E, idx, ex16, eps_tie=2.0 ** -40)
                r.update(latents=lk, codebook=ck, K=K, dtype=str(dt).replace("torch.", ""), variant=name)
                total_hard += r["hard"]
                log(r)
            flips = int((ex16 != exact[:n16]).sum())
            log({"latents": lk, "codebook": ck, "K": K, "dtype": str(dt).replace("torch.", ""),
                 "input_rounding_flips_vs_fp32_rows": flips, "rows": n16, "frac": flips / n16})
        del z, E
    log({"total_hard_mismatches": total_hard, "pass": total_hard == 0})
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")
    sys.exit(0 if total_hard == 0 else 1)


if __name__ == "__main__":
    main()
