"""Turn gpurun_out/prof/* (tools/collect_profiles.sh) into the tracked summaries under profiles/r01/:
launch list + per-kernel share, metric tables of the full captures, traffic.json for bench.py."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "prof")
DST = os.path.join(ROOT, "profiles", "r01")


def launch_summary():
    rows = [r for r in csv.reader(l for l in open(os.path.join(SRC, "launches.csv")) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    unit = hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[unit] in ("ns", "nsecond") else v
        t = tot.setdefault(r[kn], [0.0, 0])
        t[0] += v; t[1] += 1
    s = sum(t[0] for t in tot.values())
    with open(os.path.join(DST, "launches_bench_summary.txt"), "w") as f:
        for k, (v, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{v:10.1f} us {100*v/s:5.1f}%  n={n:4d} avg={v/n:8.2f} us  {k[:100]}\n")
    import shutil
    shutil.copy(os.path.join(SRC, "launches.csv"), os.path.join(DST, "launches_bench.csv"))


def metrics(rep, out):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_metrics.py"), os.path.join(SRC, rep)],
                         capture_output=True, text=True).stdout
    open(os.path.join(DST, out), "w").write(txt)
    vals = {}
    for line in txt.splitlines():
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if line.startswith(key):
                unit = line.split()[1]
                first = float(line.split("[")[1].split(",")[0].strip("' ]"))
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                vals[key] = int(first * mult)
    return vals


if __name__ == "__main__":
    launch_summary()
    t = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures in this "
                  "directory (cold-cache, serialised replays); read by bench.py for roofline.traffic"}
    v = metrics("lookup_cfg1.ncu-rep", "ncu_lookup_cfg1.txt")
    t["lookup_fwd_px_kernel<float,float,4,4,false,true>@cfg1"] = {
        "read": v["dram__bytes_read.sum"], "write": v["dram__bytes_write.sum"], "source": "ncu_lookup_cfg1.txt"}
    v = metrics("build_cfg1.ncu-rep", "ncu_build_cfg1.txt")
    t["corr_build_kernel<float,float>@cfg1"] = {
        "read": v["dram__bytes_read.sum"], "write": v["dram__bytes_write.sum"], "source": "ncu_build_cfg1.txt"}
    if os.path.exists(os.path.join(SRC, "lookup_cfg3.ncu-rep")):
        metrics("lookup_cfg3.ncu-rep", "ncu_lookup_cfg3.txt")
    json.dump(t, open(os.path.join(DST, "traffic.json"), "w"), indent=2)
    print(open(os.path.join(DST, "launches_bench_summary.txt")).read())
    print(json.dumps(t, indent=1))
