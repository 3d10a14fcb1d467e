#!/usr/bin/env python
"""Benchmark of the vector-quantizer hot path (BASELINE.json metric: gesture chunks quantized/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload tokenize|train] [--rows R] [--codes K] [--dtype f32|bf16]

Default workload = BASELINE.json configs[1]: config/VQ-VAE_GENEA.yml shapes (K=400 codes, D=400),
full-dataset tokenisation of 1M synthetic gesture-chunk latents per GPU, fp32, 1 x B200.
A "step" is one pass of the hot path over that batch.  Multi-GPU runs are launched by torchrun
(one rank per GPU); rows shard across ranks with no data-path collective (weak scaling: every
rank holds --rows rows).  `--workload train` times the EMA training step (search + gather/loss +
backward + EMA update, with the NCCL all-reduce of the packed statistics when N > 1).

One JSON line is printed by rank 0.  `value` is device-resident throughput, `e2e` the same metric
through the host-buffer entry point (H2D of the rows and D2H of the ids inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gesture chunks quantized/sec"
UNIT = "chunks/s"
D_LATENT = 400                       # hidden_size 200 x n_layers 2 (config/VQ-VAE*.yml)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tokenize", choices=["tokenize", "train", "kmeans"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="rows (chunks) per GPU per step")
    ap.add_argument("--codes", type=int, default=400, help="codebook size K (GENEA 400, Trinity 512)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# -------------------------------------------------------------------------------------------------
def cpu_reference_rate(workload: str, K: int, D: int, budget_s: float, block: int = 65536):
    """Reference quantizer (torch-CPU port of the reference modules, oracle/torch_port.py) on the
    host cores: bounded sample of the same synthetic workload, all host threads."""
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(block, D, generator=g)
    torch.manual_seed(0)
    if workload == "kmeans":
        # the reference's k-means IS sklearn (Clustering.py:718): one Lloyd iteration = max_iter=1 from a fixed init
        from sklearn.cluster import KMeans as SkKMeans
        zn = z.numpy()
        init = zn[:K].copy()

        def step():
            return SkKMeans(n_clusters=K, init=init, n_init=1, max_iter=1, tol=0.0, algorithm="lloyd").fit(zn)
    elif workload == "tokenize":
        mod = P.PortVQ(K, D, 0.25).eval()
        with torch.no_grad():
            mod._embedding.weight.normal_()

        def step():
            return P.tokenize_blocks(z, mod, block)
    else:
        mod = P.PortVQEMA(K, D, 0.25, 0.85, flavour="dae").train()
        with torch.no_grad():
            mod._embedding.weight.normal_()
        gq = torch.randn(block, D, generator=g)

        def step():
            x = z.clone().requires_grad_(True)
            loss, q, _, _ = mod(x)
            (loss + (q * gq).sum()).backward()
            return x.grad
    step()
    step()                                                    # 2 warm-ups
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 5 or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if len(times) >= 200:
            break
    times.sort()
    med = times[len(times) // 2]
    if workload == "kmeans":      # fit(max_iter=1) = one Lloyd iteration + the final assignment pass: two passes over the rows
        return dict(value=2 * block / med, unit=UNIT, cores=cores, kind="reference",
                    sample=f"{len(times)} x sklearn KMeans(max_iter=1).fit on one {block}-row block (K={K}, D={D}, fp32): "
                           f"2 passes over the rows per fit, median; sklearn {__import__('sklearn').__version__}")
    return dict(value=block / med, unit=UNIT, cores=cores, kind="port",
                sample=f"{len(times)} passes over one {block}-row block of the same synthetic workload "
                       f"(K={K}, D={D}, fp32), median; torch {torch.__version__} CPU, {torch.get_num_threads()} threads")


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, D = a.codes, D_LATENT
    block = 65536
    from oracle import torch_port as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(block, D, generator=g)
    passes = 1
    if a.workload == "kmeans":
        from sklearn.cluster import KMeans as SkKMeans
        zn, passes = z.numpy(), 2
        init = zn[:K].copy()
        fn = lambda: SkKMeans(n_clusters=K, init=init, n_init=1, max_iter=1, tol=0.0, algorithm="lloyd").fit(zn)   # noqa: E731
    elif a.workload == "tokenize":
        mod = P.PortVQ(K, D, 0.25).eval()
        with torch.no_grad():
            mod._embedding.weight.normal_()
        fn = lambda: P.tokenize_blocks(z, mod, block)       # noqa: E731
    else:
        mod = P.PortVQEMA(K, D, 0.25, 0.85, flavour="dae").train()
        gq = torch.randn(block, D, generator=g)

        def fn():
            x = z.clone().requires_grad_(True)
            loss, q, _, _ = mod(x)
            (loss + (q * gq).sum()).backward()
    for _ in range(a.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    val = passes * block * a.steps / dt
    sample = (f"each step = one {block}-row block (bounded sample of the {a.rows}-row workload), "
              f"reference quantizer ops on CPU via oracle/torch_port.py (the Python reference cannot travel)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, 1, "cpu"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(a, world, path):
    name = ("config/VQ-VAE_GENEA.yml full-dataset tokenization" if a.codes == 400 else
            f"tokenization K={a.codes}") if a.workload == "tokenize" else (
        f"k-means Lloyd iteration (assign + centre update) K={a.codes}" if a.workload == "kmeans" else
        f"VQ-VAE EMA training step (fwd+bwd+EMA) K={a.codes}")
    return {"workload": name, "codes_K": a.codes, "latent_dim_D": D_LATENT, "rows_per_gpu": a.rows,
            "rows_total": a.rows * world, "latent_dtype": a.dtype, "sharding": f"rows x{world}, codebook replicated",
            "search_path": path, "l2_policy": "inputs larger than L2 (rows*D*bytes >> 126 MB)",
            "latents": "iid N(0,1) (worst case for near-ties)" if a.workload in ("tokenize", "kmeans") else
                       "clustered: E[c] + 0.1*N(0,1), c ~ Zipf(1.1)"}


# -------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the quantizer path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import gesture2vec_b200 as g2v
    from gesture2vec_b200 import _lib
    lib = _lib.load()

    K, D, N = a.codes, D_LATENT, a.rows
    flags = {"auto": _lib.ALGO_AUTO, "simt": _lib.ALGO_SIMT, "tc": _lib.ALGO_TC}[a.algo]
    tdt = torch.float32 if a.dtype == "f32" else torch.bfloat16
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    z = torch.randn(N, D, device=dev, generator=gen).to(tdt)
    torch.manual_seed(0)
    E = torch.randn(K, D, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    pk = peaks()

    if a.workload == "tokenize":
        cb = g2v.prepare_codebook(E)
        idx = torch.empty(N, dtype=torch.int32, device=dev)

        def step():
            g2v.vq_search(z, E, cb, flags=flags, stats=stats, out=idx)
        bytes_per_row = D * z.element_size() + 4
        launches_per_step = None
    elif a.workload == "kmeans":
        # one Lloyd iteration per step: assignment (search), residual sums + counts (apply, no output rows),
        # all-reduce of the packed statistics when sharded, centre update + codebook aux refresh
        from gesture2vec_b200.kmeans import kmeans_update
        zf = z.float()
        Ek = zf[:K].clone()
        Ek2 = torch.empty_like(Ek)
        cbk = g2v.prepare_codebook(Ek)
        kidx = torch.empty(N, dtype=torch.int32, device=dev)
        red = g2v.StatsAllReduce() if world > 1 else None
        state = {"E": Ek, "E2": Ek2}

        def step():
            E0, E1 = state["E"], state["E2"]
            g2v.vq_search(z, E0, cbk, flags=flags, stats=stats, out=kidx)
            _, packed = g2v.vq_apply(zf, E0, kidx, want_out=False, want_stats=True, want_dwr=True)
            if red is not None:
                red(packed)
            kmeans_update(E0, packed, E1, None, cbk)
            state["E"], state["E2"] = E1, E0
        bytes_per_row = D * z.element_size() + 4 + D * 4 + 4      # search pass + statistics pass
        launches_per_step = None
    else:
        layer = g2v.DAE_VQ_Payam_EMA(K, D, 0.25, 0.85).to(dev).train()
        with torch.no_grad():
            layer._embedding.weight.copy_(E)
        layer.search_flags = flags
        layer.return_encodings = False
        if world > 1:
            g2v.enable_data_parallel_ema(layer)
        # training latents: clustered around the codes with Zipf-distributed usage;
        # iid noise has no cluster structure, so an EMA codebook trained on it collapses to the origin
        w = 1.0 / torch.arange(1, K + 1, device=dev, dtype=torch.float64) ** 1.1
        code = torch.multinomial(w / w.sum(), N, replacement=True, generator=gen)
        zf = (E[code] + 0.1 * z.float()).requires_grad_(True)
        gq = torch.randn(N, D, device=dev, generator=gen)

        def step():
            zf.grad = None
            loss, q, ppl, _ = layer(zf)
            torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
        bytes_per_row = 8008
        launches_per_step = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    stats.zero_()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-step CUDA events around the dominant kernel, recorded by the library on the launch stream
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for e0, e1 in kev:                       # events must exist (be created) before their handles are passed
        e0.record(); e1.record()
    barrier()
    ev0.record()
    for i in range(a.steps):
        lib.g2v_profile_next_search(kev[i][0].cuda_event, kev[i][1].cuda_event)
        step()
    ev1.record()
    barrier()
    kernel_ms = sum(e0.elapsed_time(e1) for e0, e1 in kev) / a.steps
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = float(ms.item())
    ms_step = ms_total / a.steps
    value = N * world / (ms_step * 1e-3)
    st = stats.cpu().numpy().tolist()

    # ---- end-to-end through the host-buffer entry point (tokenize only) ----
    e2e = None
    if a.workload == "tokenize" and not a.no_e2e:
        zh = torch.empty(N, D, dtype=tdt, pin_memory=True)
        zh.copy_(z)
        ih = torch.empty(N, dtype=torch.int32, pin_memory=True)
        cb = g2v.prepare_codebook(E)
        e_steps = max(3, min(a.steps, 5))
        g2v.tokenize_host(zh, E, cb, chunk_rows=131072, out=ih, flags=flags)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            g2v.tokenize_host(zh, E, cb, chunk_rows=131072, out=ih, flags=flags)   # returns after completion
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": N * world * e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": N * D * z.element_size(), "d2h_bytes_per_step": N * 4,
               "steps": e_steps, "api": "g2v_tokenize_host (pinned host rows -> host int32 ids)"}
        assert torch.equal(ih, idx.cpu()), "host path and device path disagree"
    elif a.workload == "kmeans" and not a.no_e2e:
        zh = torch.empty(N, D, dtype=torch.float32, pin_memory=True).copy_(z.float())
        init = zh[:K].numpy().copy()
        iters = 4
        kw = dict(n_clusters=K, init=init, max_iter=iters, tol=0.0, device=dev,
                  stats_reduce=g2v.StatsAllReduce() if world > 1 else None,
                  count_reduce=(lambda t: dist.all_reduce(t)) if world > 1 else None)
        g2v.KMeans(**kw).fit(zh)                 # untimed: first-use allocations
        barrier()
        t0 = time.perf_counter()
        km = g2v.KMeans(**kw).fit(zh)            # labels_ come back to the host
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": N * world * (km.n_iter_ + 1) / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": N * 4 + K * D * 4, "steps": 1,
               "api": f"KMeans.fit(host rows), {km.n_iter_} Lloyd iterations + final assignment; rows copied once"}
    elif a.workload == "train" and not a.no_e2e:
        zh = torch.empty(N, D, dtype=torch.float32, pin_memory=True).copy_(z.float())
        e_steps = 3
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            zin = zh.to(dev, non_blocking=True).requires_grad_(True)
            loss, q, ppl, _ = layer(zin)
            torch.autograd.backward([loss, q], [torch.ones_like(loss), gq])
            _ = loss.item()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": N * world * e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": 4, "steps": e_steps,
               "api": "quantizer module forward+backward, pinned host rows in, loss.item() out"}

    if rank == 0:
        path = "simt-fp32" if lib.g2v_search_path(K, D, flags) == _lib.ALGO_SIMT else "tcgen05-fp16+exact-rerank"
        flops_per_row = 2.0 * K * D
        t_hbm = bytes_per_row / (pk["hbm"] * 1e9)
        t_tc = flops_per_row / (pk["bf16_sustained"] * 1e12)
        # dominant kernel (tcgen05 sweep / fp32 sweep) timed alone by CUDA events on its launch stream
        search_bytes_per_row = D * z.element_size() + 4
        sec_per_row_kernel = (kernel_ms * 1e-3) / N
        sec_per_row_step = (ms_step * 1e-3) / N
        if t_hbm >= t_tc:
            roof = {"bound": "hbm", "achieved": search_bytes_per_row / sec_per_row_kernel / 1e9, "peak": pk["hbm"],
                    "unit": "GB/s", "whole_step_achieved": bytes_per_row / sec_per_row_step / 1e9}
        else:
            roof = {"bound": "tensor", "achieved": flops_per_row / sec_per_row_kernel / 1e12,
                    "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "whole_step_achieved": flops_per_row / sec_per_row_step / 1e12}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["whole_step_frac"] = roof["whole_step_achieved"] / roof["peak"]
        tn = TRAFFIC_NOTE.get((a.workload, K))
        # valid only for the captured configuration (same rows / dtype / kernel); scaled to this launch's rows
        roof["traffic"] = (tn["dram_bytes_per_launch"] * N / tn["rows"]
                           if tn and tn["dtype"] == a.dtype and tmem_variant(a.dtype, N, K, D) else None)
        roof["traffic_source"] = tn["source"] if roof["traffic"] is not None else None
        roof["peak_source"] = pk["src"] + (" (sustained bf16)" if roof["bound"] == "tensor" else " (copy)")
        roof["algorithmic_per_chunk"] = {"search_bytes": search_bytes_per_row, "step_bytes": bytes_per_row,
                                         "flops": flops_per_row}
        roof["kernel"] = "search_simt_kernel" if path.startswith("simt") else (
            f"tc_tmem_kernel (tcgen05 sweep, {a.dtype} rows -> fp16 operand in TMEM)" if tmem_variant(a.dtype, N, K, D)
            else "tc_search_kernel (tcgen05 sweep)")
        roof["kernel_ms_per_launch"] = kernel_ms
        roof["kernel_share_of_step"] = kernel_ms / ms_step
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic", "config": workload_config(a, world, path),
            "roofline": roof, "e2e": e2e, "clocks": clocks,
            "search_stats": {"pair_recheck_rows": st[1], "full_recheck_rows": st[2], "fallback_rows": st[3],
                             "rows": N * a.steps},
        }
        out["gpu_launches"] = int(os.environ.get("G2V_LAUNCHES_PER_STEP", "0")) * a.steps or launches_estimate(a, path) * a.steps
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference_rate(a.workload, K, D, a.cpu_seconds)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read+write per launch of the dominant kernel from the committed `ncu --set full`
# captures under profiles/ (None where no capture exists for that workload)
TRAFFIC_NOTE = {
    # tc_tmem_kernel<float>, 1M fp32 rows, K=400: 1.6005 GB read + 0.0089 GB written (profiles/r1_ncu_full_tc_tmem_k400.csv)
    # against 1.604 GB algorithmic -- every row is read from DRAM exactly once, the codebook stays in L2
    ("tokenize", 400): {"dram_bytes_per_launch": 1.6094e9, "rows": 1_000_000, "dtype": "f32",
                        "source": "profiles/r1_ncu_full_tc_tmem_k400.csv"},
}


def tmem_variant(dtype, N, K, D):
    """Mirror of plan_tmem() in csrc/g2v_tc.cu: rows readable by TMA and at most four code tiles."""
    if D % (4 if dtype == "f32" else 8) or N <= 128 or D < 64:
        return False
    dp = (D + 15) // 16 * 16
    acc0 = (dp // 2 + 15) // 16 * 16
    nt_max = min(256, ((512 - acc0) // 2) & ~15)
    return nt_max >= 32 and -(-K // nt_max) <= 4


def launches_estimate(a, path):
    """Kernels of ours launched per step (counted from the launch sites in csrc/)."""
    # (checked against the ncu launch lists under profiles/)
    # simt: fp32 sweep + per-row fp64 re-rank + batched re-rank (lists longer than 4096 rows)
    # tc:   [row_prep unless the fp32 rows are converted inside the sweep] + tcgen05 sweep + rerank_kernel
    #       (candidate / chain / whole-row lists) + batched re-rank of an overflowing whole-row list
    fused = tmem_variant(a.dtype, a.rows, a.codes, 400) or (a.dtype == "f32" and a.codes <= 512)
    search = 3 if (path.startswith("simt") or fused) else 4
    if a.workload == "tokenize":
        return search
    if a.workload == "kmeans":      # search + apply + pack + centre update + codebook prep(3)
        return search + 1 + 1 + 1 + 3
    # train: search + apply + pack + finalize + ema(2) + codebook prep(3) + backward
    return search + 1 + 1 + 1 + 2 + 3 + 1


if __name__ == "__main__":
    main()
