"""Turn an Nsight Compute report of the fused kernel into the committed JSON summaries.
Usage: python scripts/ncu_summarize.py gpurun_out/r02_prof.ncu-rep profiles/r02_ncu_full_configB_metrics.json [config] [n_gpus]
Writes the per-launch values of the metrics that matter for the roofline block and refreshes profiles/ncu_summary.json
(`dram_bytes_per_launch`, read by bench.py for `roofline.traffic` at the matching config / GPU count)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_bytes.sum", "smsp__cycles_active.avg", "launch__occupancy_limit_shared_mem",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cfg = sys.argv[3] if len(sys.argv) > 3 else "B"
    n_gpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    res = {}
    for name in KEEP:
        if name in header:
            i = header.index(name)
            res[name] = {"unit": units[i], "per_launch": [r[i] for r in data]}
    kn = header.index("Kernel Name") if "Kernel Name" in header else None
    res["kernels"] = sorted({r[kn] for r in data}) if kn is not None else []
    json.dump(res, open(out, "w"), indent=1)

    def to_bytes(v, unit):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    if "dram__bytes_read.sum" in res and "dram__bytes_write.sum" in res:
        rd, wr = res["dram__bytes_read.sum"], res["dram__bytes_write.sum"]
        n = len(rd["per_launch"])
        total = sum(to_bytes(rd["per_launch"][i], rd["unit"]) + to_bytes(wr["per_launch"][i], wr["unit"]) for i in range(n)) / n
        json.dump({"source": f"{out} (ncu --set full --clock-control none, fm_moe_forward_kernel, config {cfg}, {n_gpus} GPU)",
                   "config": cfg, "n_gpus": n_gpus, "dram_bytes_per_launch": total}, open("profiles/ncu_summary.json", "w"))
        print(f"dram bytes per launch {total / 1e6:.1f} MB over {n} launches")
    for k in ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
              "launch__registers_per_thread", "l1tex__m_xbar2l1tex_read_bytes.sum"):
        if k in res:
            print(k, res[k]["unit"], res[k]["per_launch"])


if __name__ == "__main__":
    main()
