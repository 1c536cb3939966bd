"""Turn the raw ncu outputs in gpurun_out/ into small committed summaries under profiles/."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

KNOWN = ("gemm_bf16_kernel", "attention_tc_kernel", "attention_gated_kernel", "posconv_kernel", "layernorm_kernel", "fbank_kernel",
         "patchify_kernel", "group_pad_kernel", "mean_pool_kernel", "f32_to_bf16", "posconv_pack", "posconv_norm", "gate_pack",
         "melspec_kernel", "melspec_normalise", "melspec_init", "dwconv_tma_kernel", "dwconv_kernel", "pointwise_kernel", "se_mlp_kernel",
         "se_apply_kernel", "stem_kernel", "nhwc_to_nchw_kernel")


def launches(csv_path, out_name, span_marker=None):
    """Per-kernel launch list: count, device time and share; DRAM bytes per launch when the capture holds them
    (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)."""
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr_i]
    kn, mn, mv, mu, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    per = collections.OrderedDict()  # launch id -> {name, ms, rd, wr}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[hdr_i + 1:]:
        v = float(r[mv].replace(",", ""))
        unit = r[mu]
        name = r[kn].split("(")[0]
        for k in KNOWN:
            if k in name:
                name = k
        # template arguments distinguish the GEMM epilogues: keep mode / pair / warps
        if name == "dwconv_tma_kernel" and "<" in r[kn]:
            name += "<" + r[kn].split("<", 1)[-1].split(">")[0].replace("(int)", "").replace(", ", " ") + ">"  # <K S CH>
        if name == "gemm_bf16_kernel" and "<" in r[kn]:
            name += "<" + r[kn].split("<", 2)[-1].split(">")[0].replace("(int)", "").replace("(bool)", "") + ">"
        e = per.setdefault(r[idc], {"name": name, "ms": 0.0, "rd": None, "wr": None})
        if r[mn].startswith("gpu__time_duration"):
            e["ms"] = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v if unit.startswith("ms") else v * 1e3
        elif r[mn].startswith("dram__bytes_read"):
            e["rd"] = v * scale.get(unit, 1.0)
        elif r[mn].startswith("dram__bytes_write"):
            e["wr"] = v * scale.get(unit, 1.0)
    if span_marker:  # keep ONE complete pass: from the first launch of `span_marker` to the one before its next occurrence
        ids = list(per)
        marks = [i for i, k in enumerate(ids) if per[k]["name"].startswith(span_marker)]
        if len(marks) >= 2:
            per = collections.OrderedDict((k, per[k]) for k in ids[marks[0]:marks[1]])
    agg = collections.OrderedDict()
    total = 0.0
    for e in per.values():
        a = agg.setdefault(e["name"], [0, 0.0, 0.0, 0.0, False])
        a[0] += 1; a[1] += e["ms"]; total += e["ms"]
        if e["rd"] is not None:
            a[2] += e["rd"]; a[3] += e["wr"] or 0.0; a[4] = True
    lines = ["# ncu --clock-control none launch list (cold-cache, serialised: compare SHARES, not absolutes)",
             f"# source: {os.path.basename(csv_path)}; total {total:.3f} ms over {len(per)} launches",
             "kernel,launches,total_ms,share,avg_ms,avg_dram_read_MB,avg_dram_write_MB"]
    for k, (n, ms, rd, wr, has) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k},{n},{ms:.4f},{ms / total:.4f},{ms / n:.4f}," + (f"{rd / n / 1e6:.2f},{wr / n / 1e6:.2f}" if has else ","))
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:16]))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
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
    if os.path.exists(os.path.join(g, f"launches_effnet_{tag}.csv")):  # one complete EfficientNet forward
        launches(os.path.join(g, f"launches_effnet_{tag}.csv"), f"launches_effnet_{tag}.csv", span_marker="melspec_init")
    for rep in sys.argv[2:]:
        full(os.path.join(g, rep + ".ncu-rep"), rep + "_summary.json")
