"""Summarise an ncu report (.ncu-rep) of the render kernel into a small JSON + markdown pair under profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_fine_vanilla_f16x3 [launch_index]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "lts__t_bytes.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    vals = rows[2 + idx]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or ("tensor" in h and "pct_of_peak_sustained_elapsed" in h and ".avg" in h) or h == "Kernel Name":
            d[h] = {"value": v, "unit": u}

    def num(k):
        v, u = float(d[k]["value"].replace(",", "")), d[k]["unit"].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    summary = {"report": rep, "launch_index": idx, "metrics": d,
               "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum")}
    json.dump(summary, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("# ncu --set full summary: %s (launch %d)\n\n" % (rep, idx))
        for k, v in d.items():
            if "ops_path_tensor" in k and v["value"].strip() in ("0", "0.0"):
                continue                 # the per-format tensor paths this kernel does not use
            f.write("- `%s` = %s %s\n" % (k, v["value"], v["unit"]))
        f.write("\ndram bytes per launch (read+write) = %.0f\n" % summary["dram_bytes_per_launch"])
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
