"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
bench.py: find one training step (from a log-mel / first-kernel marker to the next), group by kernel family, print
time / DRAM bytes / GB/s per family, and optionally emit the traffic record bench_workloads._committed_traffic reads.
usage: python tools/launch_list_summary.py launches.csv <launches_per_step> [marker-substring] [json-key out.json]"""
import collections, csv, json, re, sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
per_step = int(sys.argv[2])
marker = sys.argv[3] if len(sys.argv) > 3 else "logmel_kernel"
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
L = list(launch.values())
starts = [i for i, l in enumerate(L) if marker in l["name"]]
i0 = starts[0]
if i0 + per_step <= len(L):
    step = L[i0:i0 + per_step]
else:                                   # the window holds the tail of one step and the head of the next
    step = L[i0:] + L[i0 - per_step + (len(L) - i0) - 0:i0][: per_step - (len(L) - i0)]
    step = L[i0:] + L[len(L) - per_step:i0]
assert len(step) == per_step, (len(step), per_step, starts, len(L))

def family(n):
    m = re.match(r"(?:void )?([\w:]+)", n).group(1)
    t = re.search(r"<(.*?)>", n)
    if "gemm_tc_kernel" in m and t:
        return "tc::gemm_tc_kernel<" + ("bf16" if "bfloat16" in t.group(1).split(",")[0] else "tf32") + ">"
    return m

fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in step:
    f = fam[family(l["name"])]
    f[0] += 1
    f[1] += l["gpu__time_duration.sum"] / 1e3
    f[2] += l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)
    f[3] += l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * l["gpu__time_duration.sum"] / 1e3
tot = sum(f[1] for f in fam.values()); totb = sum(f[2] for f in fam.values())
print(f"one step: {per_step} launches, {tot:.1f} us serialised under ncu (cold caches), {totb/1e9:.3f} GB DRAM traffic")
print(f"{'kernel family':46s} {'n':>4s} {'us':>9s} {'share':>6s} {'DRAM MB':>9s} {'GB/s':>7s} {'us/launch':>9s} {'tensor%':>7s}")
for k, f in sorted(fam.items(), key=lambda x: -x[1][1]):
    print(f"{k:46s} {f[0]:4d} {f[1]:9.1f} {100*f[1]/tot:5.1f}% {f[2]/1e6:9.1f} {f[2]/max(f[1],1e-9)/1e3:7.0f} {f[1]/f[0]:9.1f} {f[3]/max(f[1],1e-9):7.1f}")
if len(sys.argv) > 5:
    key, out = sys.argv[4], sys.argv[5]
    opmap = {"tc::gemm_tc_kernel<bf16>": "lr_gemm_bf16", "tc::gemm_tc_kernel<tf32>": "lr_gemm_tf32", "bn::bn_act_fwd_kernel": "lr_bn_act_fwd_h",
             "ig::conv3x3_kernel": "lr_conv3x3_bf16", "ig::conv3x3_wgrad_kernel": "lr_conv3x3_wgrad_bf16"}
    fams = {}
    for k, f in fam.items():
        if k in opmap: fams[opmap[k]] = {"dram_bytes": f[2], "launches": f[0]}
    bwd = [f for k, f in fam.items() if "bn_act_bwd" in k]
    if bwd: fams["lr_bn_act_bwd_h"] = {"dram_bytes": sum(f[2] for f in bwd), "launches": max(f[0] for f in bwd)}
    rec = json.load(open(out)) if __import__("os").path.exists(out) else {}
    rec[key] = {"families": fams, "step_dram_bytes": totb, "source": sys.argv[1].split("/")[-1]}
    json.dump(rec, open(out, "w"), indent=1)
