"""Turn the raw ncu outputs in gpurun_out/ into small committed summaries under profiles/."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

def launches(csv_path, out_name):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr_i]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[hdr_i + 1:]:
        v = float(r[mv].replace(",", ""))
        unit = r[mu]
        ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v if unit.startswith("ms") else v * 1e3
        name = r[kn].split("(")[0]
        for k in ("gemm_bf16_kernel", "attention_gated_kernel", "posconv_kernel", "layernorm_kernel", "fbank_kernel", "patchify_kernel",
                  "group_pad_kernel", "mean_pool_kernel", "f32_to_bf16", "posconv_pack", "posconv_norm", "gate_pack"):
            if k in name:
                name = k
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms; total += ms
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)", f"# source: {os.path.basename(csv_path)}; total {total:.3f} ms over {sum(a[0] for a in agg.values())} launches", "kernel,launches,total_ms,share"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k},{n},{ms:.4f},{ms / total:.4f}")
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]

def full(rep, out_name):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    kn = h.index("Kernel Name")
    cols = [i for i, c in enumerate(h) if c in WANT]
    out = []
    for d in data:
        rec = {"kernel": d[kn].split("(")[0][-60:]}
        for i in cols:
            rec[h[i] + (" [" + units[i] + "]" if units[i] else "")] = d[i]
        out.append(rec)
    json.dump(out, open(os.path.join(OUT, out_name), "w"), indent=1)
    for r in out:
        print(r["kernel"], {k.split(".")[0][-28:]: v for k, v in list(r.items())[1:8]})

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    if os.path.exists(os.path.join(g, f"launches_{tag}.csv")):
        launches(os.path.join(g, f"launches_{tag}.csv"), f"launches_{tag}.csv")
    for rep in sys.argv[2:]:
        full(os.path.join(g, rep + ".ncu-rep"), rep + "_summary.json")
