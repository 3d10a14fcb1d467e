"""Markdown summary of a bench.py default-run JSON line (used for profiles/README.md)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"* headline ({d['n_gpus']} GPU): {d['value']:.3e} chunks/s, {d['ms_per_step']:.3f} ms/step; `{r['kernel']}` {r['kernel_ms_per_launch']:.3f} ms "
      f"= {r['achieved']:.0f} {r['unit']} of {r['peak']:.0f} (frac {r['frac']:.3f}), whole step {r['whole_step_frac']:.3f}; "
      f"traffic {r['traffic'] / 1e9 if r['traffic'] else float('nan'):.3f} GB; clocks {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
if d.get("sustained"):
    print(f"* sustained (>= 1 s): {d['sustained']['ms_per_step']:.3f} ms/step")
if d.get("e2e"):
    print(f"* e2e: {d['e2e']['value']:.3e} chunks/s ({d['e2e'].get('h2d_gbs_per_gpu', 0):.1f} GB/s H2D per GPU)")
if d.get("cpu_baseline"):
    print(f"* cpu_baseline: {d['cpu_baseline']['value']:.3e} chunks/s on {d['cpu_baseline']['cores']} threads ({d['cpu_baseline']['kind']})")
for t in d.get("train", []) or []:
    for k in ("bulk", "dropin_onehot", "bulk_stream_ordered", "dropin_onehot_stream_ordered"):
        if k in t:
            v = t[k]
            print(f"* train K={t['codes_K']} {k}: {v['ms_per_step']:.3f} ms/step, {v['value']:.3e} chunks/s, frac {v['roofline']['frac']:.3f}, "
                  f"launches {v['own_launches_per_step']}" + (f", overlap_frac {v.get('overlap_frac')}" if 'overlap_frac' in v else ""))
    if "allreduce_us" in t:
        print(f"  all-reduce alone: {t['allreduce_us']:.1f} us")
if d.get("dp_check"):
    print("* dp_check:", d["dp_check"])
print("| K | rows | ms/step | TFLOP/s step | kernel ms | frac of sustained bf16 (step / kernel) | re-rank rows |")
print("|---|---|---|---|---|---|---|")
for s in d.get("sweep", []) or []:
    print(f"| {s['codes_K']} | {s['latent_dtype']} | {s['ms_per_step']:.3f} | {s['tflops_step']:.0f} | {s['sweep_kernel_ms']:.3f} | "
          f"{s['frac_of_sustained_bf16_peak_step']:.2f} / {s['frac_of_sustained_bf16_peak_kernel']:.2f} | {100 * s['rerank_rows_frac']:.1f} % |")
for k in ("latency_n128_us", "soft_quantizer", "vqvae_ema_flavour", "eager_cuda_baseline", "vs_eager_cuda"):
    if d.get(k):
        print(f"* {k}: {json.dumps(d[k])}")
