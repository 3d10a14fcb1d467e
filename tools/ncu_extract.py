"""Summarise Nsight Compute output for profiles/ (run in the authoring container, after gpurun).

    python tools/ncu_extract.py full  <report.ncu-rep> <out.csv>      # key metrics of every launch in a --set full report
    python tools/ncu_extract.py list  <launches.csv>   <out.csv>      # compact per-launch list (kernel, grid, block, us)
"""
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    return name.replace("g2v::<unnamed>::", "").replace("void ", "").strip()


def full(rep: str, out: str) -> None:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}:{short(r[col['Kernel Name']])}" for i, r in enumerate(launches)])
        for k in KEYS:
            if k in col:
                w.writerow([k, units[col[k]]] + [r[col[k]] for r in launches])


def launch_list(src: str, out: str) -> None:
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "gpu__time_duration_us"])
        for r in rows:
            w.writerow([r[0], short(r[4])[:90], r[7], r[8], f"{float(r[-1]) / 1e3:.1f}"])


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
