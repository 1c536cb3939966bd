"""Attribute ncu warp-stall samples to CUDA source lines: joins `ncu --page source --csv` (SASS view) with the line table
of the kernel's cubin (`nvdisasm -g`).   python tools/ncu_lines.py report.ncu-rep cubin kernel-substring [launch-id] [top]"""
import csv, re, subprocess, sys
from collections import defaultdict


def sass_rows(rep, kid):
    args = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if kid is not None:
        args += ["--kernel-id", f":::{kid}"]
    out = subprocess.run(args, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
    hdr = rows[hi]
    data, seen = [], set()
    for r in rows[hi + 1:]:
        if len(r) > 3 and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    return hdr, data


def line_table(cubin, func_sub):
    out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    table, cur, active = {}, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
        if m:
            active = func_sub in m.group(1)
            continue
        if not active:
            continue
        if "//## File" in ln:
            # attribute inlined helpers (ptx.cuh wrappers ...) to their outermost call site
            ms = re.findall(r'"([^"]+)", line (\d+)', ln)
            if ms:
                f, l = ms[-1]
                cur = (f.split("/")[-1], int(l), len(ms) > 1)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, cubin, sub = sys.argv[1], sys.argv[2], sys.argv[3]
    kid = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] not in ("", "-") else None
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    hdr, data = sass_rows(rep, kid)
    si = hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][0], 16)
    lt = line_table(cubin, sub)
    agg, why = defaultdict(int), defaultdict(lambda: defaultdict(int))
    tot = 0
    for r in data:
        n = int(r[si] or 0)
        tot += n
        key = lt.get(int(r[0], 16) - base, ("?", 0, False))[:2]
        agg[key] += n
        for i in stall:
            v = int(r[i] or 0)
            if v:
                why[key][hdr[i][6:]] += v
    print("total samples", tot)
    src = {}
    for (f, l), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        if f not in src:
            try:
                src[f] = open("/root/repo/avex_b200/csrc/" + f).read().splitlines()
            except OSError:
                src[f] = []
        text = src[f][l - 1].strip()[:80] if 0 < l <= len(src[f]) else ""
        w = sorted(why[(f, l)].items(), key=lambda kv: -kv[1])[:2]
        print(f"{n:7d} {100.0 * n / tot:5.1f}%  {f}:{l:<4d} {text}   {w}")


if __name__ == "__main__":
    main()
