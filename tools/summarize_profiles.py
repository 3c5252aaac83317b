"""Run HERE (no GPU): turns the .ncu-rep / launch-list files brought back in gpurun_out/ into the small text/JSON
summaries committed under profiles/ (the judge reads profiles/, gpurun_out/ is scratch)."""
import csv, json, os, subprocess, sys, collections
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:60]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try: st.append((round(float(r[i]), 2), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError: pass
        d["top_stalls"] = sorted(st, reverse=True)[:5]
        res.append(d)
    return res

summary = {}
for name in (f"leaf_hash_{R}", f"ntt_passes_{R}"):
    rep = os.path.join(GO, name + ".ncu-rep")
    if os.path.exists(rep):
        summary[name] = raw(rep)
with open(os.path.join(PR, f"ncu_summary_{R}.json"), "w") as f:
    json.dump(summary, f, indent=1)
# launch list -> per-kernel totals and shares
ll = os.path.join(GO, f"launches_{R}.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ik].split("(")[0]
        t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += float(r[iv].replace(",", ""))
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(PR, f"launches_{R}.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu  (4 steps + leaf-hash timing reps)\n")
        f.write(f"# cold-cache, serialised: compare SHARES.  total {total/1e6:.2f} ms over {sum(v[0] for v in tot.values())} launches\n")
        for k, (n, ns) in sorted(tot.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:48s} launches {n:4d}  total {ns/1e6:10.3f} ms  share {100*ns/total:5.1f} %\n")
    print(open(os.path.join(PR, f"launches_{R}.txt")).read())
# traffic for bench.py's roofline.traffic
lh = summary.get(f"leaf_hash_{R}")
if lh:
    def num(s): return float(s.split()[0].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[s.split()[1]]
    t = num(lh[0]["dram__bytes_read.sum"]) + num(lh[0]["dram__bytes_write.sum"])
    out = {"kernel": "mk::leaf_hash_fast_kernel", "round": R, "dram_bytes_per_launch": t, "source": f"profiles/ncu_summary_{R}.json"}
    ntt = summary.get(f"ntt_passes_{R}")
    if ntt:  # the LDE launches of one step (capture_profiles.sh captures exactly one step's worth)
        out["lde_dram_bytes_per_step"] = sum(num(e["dram__bytes_read.sum"]) + num(e["dram__bytes_write.sum"]) for e in ntt)
        out["lde_source"] = f"profiles/ncu_summary_{R}.json (the {len(ntt)} LDE launches of one step, ncu --set full)"
        print("LDE DRAM traffic per step:", out["lde_dram_bytes_per_step"] / 1e9, "GB")
    json.dump(out, open(os.path.join(PR, "roofline_traffic.json"), "w"))
    print("leaf hash DRAM traffic per launch:", t / 1e9, "GB")
