"""Per-kernel time breakdown of one bench step with torch.profiler (CUPTI) -- cheap compared with an
ncu pass over ~10^4 launches.  Scratch tool; writes gpurun_out/step_profile.txt."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

def main():
    args = types.SimpleNamespace(gpus=1, steps=1, warmup=2, impl="ours", workload=os.environ.get("WL", "v16"),
                                 fp32=False, no_cpu_baseline=True, no_eager_baseline=True, phases=False)
    os.environ["U2PL_BENCH_FAST"] = "1"
    from torch.profiler import profile, ProfilerActivity
    # warm-up outside the profiler
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.our_arm(args)          # builds + warms cuDNN autotune caches
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        with contextlib.redirect_stdout(buf):
            bench.our_arm(args)
    rows = []
    for e in prof.key_averages():
        if e.device_time_total > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
            rows.append((e.device_time_total, e.count, e.key))
    if not rows:
        for e in prof.key_averages():
            if getattr(e, "self_device_time_total", 0) > 0:
                rows.append((e.self_device_time_total, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    out = [f"total device time {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} launches (3 steps: 2 warm-up + 1)"]
    for t, n, k in rows[:60]:
        out.append(f"{t/1e3:10.2f} ms {100*t/tot:5.1f}%  n={n:6d}  {k[:140]}")
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/step_profile.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:45]))

if __name__ == "__main__":
    main()
