"""Turns gpurun_out/{launches_R.csv, prof_R_<kernel>.ncu-rep} into the tracked summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
# optional second argument: "1080p8" for the captures of scripts/gpu_prof_1080p8.sh
TAG = sys.argv[2] if len(sys.argv) > 2 else ""
WL, WLNAME, WLDESC = (("1080p8", "1080p8_n7", "1080p 8-bit 4:2:0, N=7") if TAG == "1080p8" else ("4k10", "4k10_n15", "4K 10-bit 4:2:0, N=15"))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)
# 1. launch list
src = os.path.join(root, "gpurun_out", f"launches_{R}.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        n = row["Kernel Name"].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"launches_{R}_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 (4k10_n15)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:70s} launches={v[0]:4d} total_ms={v[1]/1e6:10.3f} share={v[1]/tot*100:5.1f}%\n")
    open(os.path.join(out, f"launches_{R}.csv"), "w").writelines(lines)
# 2. per-kernel raw metrics
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
traffic = {}
for kname in ("tf_search32", "tf_search16", "tf_filter"):
    rep = os.path.join(root, "gpurun_out", f"prof_{R}_{TAG + '_' if TAG else ''}{kname}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    with open(os.path.join(out, f"ncu_{R}_{kname}_{WL}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kname}  ({WLDESC}, one launch)\n")
        for k in keys:
            if k in d:
                f.write(f"{k} = {d[k][0]} {d[k][1]}\n")
        cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
        open("/tmp/_cs.csv", "w").write(cs)
        f.write("\n# by source function (instructions / stall samples)\n")
        f.write(subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_by_function.py"), "/tmp/_cs.csv",
                                os.environ.get("TF_KERNELS_SRC", os.path.join(root, "aom-av1-psy_b200/csrc/tf_kernels.cuh")), "12"], capture_output=True, text=True).stdout)
        s = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(s.splitlines()))
        sh = srows[1]
        tot = {}
        for r in srows[2:]:
            for h, v in zip(sh, r):
                if h.startswith("stall") and "Not Issued" not in h:
                    try:
                        tot[h] = tot.get(h, 0) + int(v)
                    except ValueError:
                        pass
        ssum = sum(tot.values()) or 1
        f.write("\n# warp stall reasons (all samples)\n" + ", ".join(f"{k[6:]} {v/ssum*100:.1f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]) + "\n")
    def num(x):
        v, u = d[x]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic[kname + "_kernel"] = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
if traffic:
    p = os.path.join(out, f"traffic_{R}.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur.setdefault(WLNAME, {}).update(traffic)
    cur["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full (per-frame launches for the search kernels)"
    json.dump(cur, open(p, "w"), indent=1)
print(open(os.path.join(out, f"launches_{R}_summary.txt")).read() if os.path.exists(os.path.join(out, f"launches_{R}_summary.txt")) else "")
print(traffic)
