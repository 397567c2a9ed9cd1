"""In-graph timeline of one train step: kernel start / end times per stream from CUPTI activity records
(torch.profiler), which ncu cannot give (it serialises the launches).  Answers what the serial launch lists cannot:
how long the main chain is, how much of the weight-gradient branch is hidden under it, how long the tail after the
last main-chain kernel is and which kernels sit in it.

    python tools/step_timeline.py [--model mid_fusion_fast] [--batch 32] [--out gpurun_out/timeline]

Writes <out>.json (one record per kernel of the profiled step: name, stream, start us, duration us) and <out>.txt
(the summary printed to stdout).  A measurement aid, not a bench: numbers taken under the profiler are not bench values
(CUPTI adds ~1 us per kernel)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mid_fusion_fast")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=88)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline"))
    args = ap.parse_args()

    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench_workloads as BW

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cfg = {"num_classes": 40, "size": args.size, "grayscale": True, "cpu_batch": 8, "model": args.model,
           "precision": args.precision}
    wl = BW.AvTrainWorkload(dev, args.batch, cfg, 0, 1)
    for _ in range(8):
        wl.step_device()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:       # ONE step (one graph replay)
        wl.step_device()
        torch.cuda.synchronize()
    # the chrome trace carries the stream id of every device activity
    trace = args.out + "_trace.json"
    os.makedirs(os.path.dirname(trace), exist_ok=True)
    prof.export_chrome_trace(trace)
    tr = json.load(open(trace))
    evs = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    evs.sort(key=lambda e: e["ts"])
    os.remove(trace)
    step = evs
    t0 = min(e["ts"] for e in step)
    recs = [{"name": e["name"][:100], "stream": e["args"].get("stream"), "t0": round(e["ts"] - t0, 2), "dur": round(e["dur"], 2),
             "cat": e["cat"]} for e in step]
    end = max(r["t0"] + r["dur"] for r in recs)
    streams = {}
    for r in recs:
        streams.setdefault(r["stream"], []).append(r)
    lines = [f"model {args.model} batch {args.batch} {args.size}px {args.precision}: {len(recs)} device activities in the step, "
             f"{end:.1f} us from the first start to the last end (under CUPTI)"]
    # (the replay runs the graph's branches on internal streams; their ids say which kernels shared a branch)
    for sid, rs in sorted(streams.items(), key=lambda kv: -len(kv[1])):
        busy = sum(r["dur"] for r in rs)
        lines.append(f"  stream {sid}: {len(rs)} activities, busy {busy:.1f} us, first start {min(r['t0'] for r in rs):.1f}, "
                     f"last end {max(r['t0'] + r['dur'] for r in rs):.1f}")
    # concurrency profile: how long 0, 1, 2, 3+ kernels were running
    pts = sorted([(r["t0"], 1) for r in recs] + [(r["t0"] + r["dur"], -1) for r in recs])
    level, last, hist = 0, 0.0, {}
    for t, d in pts:
        hist[min(level, 3)] = hist.get(min(level, 3), 0.0) + (t - last)
        level, last = level + d, t
    lines.append("  time with n kernels running: " + ", ".join(f"n={k}{'+' if k == 3 else ''}: {hist.get(k, 0.0):.1f} us" for k in range(4))
                 + f"; sum of kernel durations {sum(r['dur'] for r in recs):.1f} us")
    # where nothing runs: the largest idle intervals
    idle, level, last = [], 0, 0.0
    for t, d in pts:
        if level == 0 and t - last > 0:
            idle.append((t - last, last))
        level, last = level + d, t
    idle.sort(reverse=True)
    lines.append("  largest idle intervals (us at us): " + ", ".join(f"{g:.1f}@{at:.0f}" for g, at in idle[:8])
                 + f"; idle intervals below 5 us: {sum(g for g, _ in idle if g < 5):.1f} us in {sum(1 for g, _ in idle if g < 5)}")
    lines.append("  the last 24 activities of the step (start us, duration us, stream, kernel):")
    for r in sorted(recs, key=lambda r: r["t0"])[-24:]:
        lines.append(f"    {r['t0']:8.1f} +{r['dur']:6.1f}  s{r['stream']}  {r['name'][:84]}")
    # per kernel family, in-graph durations
    fam = {}
    for r in recs:
        n = r["name"].split("<")[0].split("(")[0].replace("void ", "")
        f = fam.setdefault(n, [0, 0.0, 0.0])
        f[0] += 1
        f[1] += r["dur"]
        f[2] = max(f[2], r["dur"])
    lines.append("  in-graph time by kernel family (launches, sum us, mean us, max us):")
    for n, (cnt, us, mx) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append(f"    {n[:56]:56s} {cnt:4d} {us:8.1f} {us / cnt:7.1f} {mx:7.1f}")
    text = "\n".join(lines)
    print(text)
    with open(args.out + ".txt", "w") as f:
        f.write(text + "\n")
    with open(args.out + ".json", "w") as f:
        json.dump(recs, f)


if __name__ == "__main__":
    main()
