"""Summarise ncu artefacts into small text files for profiles/ (run here, no GPU needed).
usage: summarize_ncu.py launches <launches.csv> <out.md>
       summarize_ncu.py kernel <file.ncu-rep> <out.md> [traffic.json]"""
import csv
import io
import json
import subprocess
import sys


def launches(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    agg = {}
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        short = name.split("(")[0].replace("void ", "").replace("tnb::", "")
        if "contract_kernel" in short:
            short = short[:110]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, n, ns * 1e-6, 100 * ns / total))
        f.write("\ntotal %.3f ms over %d launches\n" % (total * 1e-6, sum(n for n, _ in agg.values())))


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__sass_l1tex_t_requests_pipe_lsu_mem_global_op_ldgsts.sum", "sm__cycles_elapsed.max"]


def kernel(rep, out, traffic_json=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary (%s)\n\n" % rep.split("/")[-1])
        for r in data:
            f.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % r[col["Kernel Name"]][:160])
            for k in KEYS:
                if k in col:
                    f.write("| %s | %s | %s |\n" % (k, r[col[k]], units[col[k]]))
            f.write("\n")
    if traffic_json:
        r = data[0]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        json.dump({"contract_kernel_dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
                   "kernel": r[col["Kernel Name"]][:200], "source": rep.split("/")[-1]}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
