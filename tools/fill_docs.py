"""Fill the @PLACEHOLDER@ numbers of DESIGN.md / README.md / profiles/README.md from the committed evidence under profiles/:
python tools/fill_docs.py   (idempotent once the placeholders are gone)."""
import csv, json, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(path):
    for l in open(path):
        if l.startswith("{"):
            return json.loads(l)
    raise SystemExit(f"no JSON line in {path}")


eff = line(os.path.join(P, "bench_effnet_r2_final.json"))
vals = {
    "EFF_MS": f"{eff['ms_per_step']:.2f}",
    "EFF_CLIPS": f"{eff['value'] / 1e3:.1f}",
    "EFF_AH": f"{eff['value'] * 5 / 3600:.1f}",
    "EFF_E2E": f"{eff['e2e']['value'] / 1e3:.1f}",
    "EFF_GBS": f"{eff['roofline']['achieved'] / 1e3:.2f}",
    "EFF_FRAC": f"{eff['roofline']['frac'] * 100:.0f}",
    "MEL_MS": f"{eff['kernels']['melspec']['ms_per_step']:.2f}",
}
fam = {"DW_MS": "dwconv_tma_kernel", "PW_MS": "pointwise_kernel"}
per_fwd = {}  # one complete forward (tools/summarize_profiles.py keeps exactly one): kernel -> total ms
for l in open(os.path.join(P, "launches_effnet_r2_final.csv")):
    if l.startswith("#") or l.startswith("kernel,"):
        continue
    parts = l.rstrip("\n").rsplit(",", 6)  # kernel names may hold commas
    per_fwd[parts[0]] = per_fwd.get(parts[0], 0.0) + float(parts[2])
for key, name in fam.items():
    vals[key] = f"{sum(v for k, v in per_fwd.items() if k.startswith(name)):.1f}"
if os.path.exists(os.path.join(P, "bench_r2_final.json")):
    b = line(os.path.join(P, "bench_r2_final.json"))
    vals.update({"B_MS": f"{b['ms_per_step']:.2f}", "B_AH": f"{b['value']:.2f}", "B_E2E": f"{b['e2e']['value']:.2f}",
                 "B_FRAC": f"{b['roofline']['frac']:.3f}", "B_MHZ": f"{b['clocks']['sm_mhz']:.0f}",
                 "B_TF": f"{b['roofline']['achieved']:.0f}"})
print(vals)
for name in ("DESIGN.md", "README.md", os.path.join("profiles", "README.md")):
    p = os.path.join(ROOT, name)
    s = open(p).read()
    t = re.sub(r"@([A-Z0-9_]+)@", lambda m: vals.get(m.group(1), m.group(0)), s)
    left = set(re.findall(r"@([A-Z0-9_]+)@", t))
    if left:
        print(f"{name}: unfilled {sorted(left)}")
    if t != s:
        open(p, "w").write(t)
        print(f"{name}: updated")
