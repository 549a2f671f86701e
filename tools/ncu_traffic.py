"""Write the DRAM-traffic sidecar that bench.py's `roofline.traffic` reads, from an `ncu --set full` capture of the chain
kernel (run here, in the build container: the .ncu-rep comes back from the GPU box in gpurun_out/).

    python tools/ncu_traffic.py gpurun_out/r02_chain_full.ncu-rep profiles/r02_chain_traffic.json [--batch 256]

Also prints the metrics quoted in profiles/*_chain_full_summary.txt."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__cluster_max_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return float(v.replace(",", "")) * mult


def main():
    rep, out = sys.argv[1], sys.argv[2]
    batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 256
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    recs = [r for r in rows[2:] if len(r) == len(hdr)]
    name_i = hdr.index("Kernel Name")
    chain = [r for r in recs if "chain_kernel" in r[name_i]]
    if not chain:
        sys.exit("no chain_kernel launch in %s" % rep)
    r = chain[0]
    vals = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals[k] = (r[i], units[i])
            print("%-70s %s %s" % (k, r[i], units[i]))
    rd, wr = to_bytes(*vals["dram__bytes_read.sum"]), to_bytes(*vals["dram__bytes_write.sum"])
    side = {"kernel": r[name_i], "batch": batch, "precision": "bf16", "dram_bytes_read": rd, "dram_bytes_write": wr,
            "dram_bytes_per_launch": rd + wr, "gpu_time_duration": " ".join(vals["gpu__time_duration.sum"]),
            "source": "ncu --set full --clock-control none of `python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary`, %s" % rep}
    json.dump(side, open(out, "w"), indent=1)
    print("wrote", out, "dram bytes per launch", rd + wr)


if __name__ == "__main__":
    main()
