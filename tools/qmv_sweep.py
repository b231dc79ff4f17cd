"""Sweep the decode GEMV work split (warps per tile, global k slices, CTAs/SM cap, cp.async stages) per linear shape of a
workload. Timing: back-to-back launches over all layers' weights (distinct weights per launch, larger than L2 for 8B models),
CUDA events around `iters` passes. Usage: python tools/qmv_sweep.py llama3-8b-int4 [iters]"""
import itertools, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from uzu_b200 import binding as B

def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b-int4"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    mdir = bench.model_dir_for(wl)
    ctx = B.Context()
    eng = B.Engine(ctx, mdir, max_context_length=512)
    names = {1: "mixer_in", 2: "mixer_out", 4: "up", 32: "up+gated", 8: "down", 16: "readout"}
    out = {}
    for sel, name in names.items():
        rows = []
        for stages, per_sm, wpt, dks in itertools.product((0, 2, 3, 4), (0,), (0, 1, 2, 4), (0, 1, 2, 3, 4, 6)):
            if sel == 32 and dks not in (0,):
                continue
            ctx.lib.uzu_debug_set_qmv_tuning(wpt, dks, per_sm, stages)
            try:
                t, n = eng.time_linears(iters, sel)
            except Exception as ex:
                print(name, stages, wpt, dks, "ERR", ex); continue
            rows.append((t / max(n, 1) * 1e6, stages, wpt, dks))
        ctx.lib.uzu_debug_set_qmv_tuning(0, 0, 0, 0)
        rows.sort()
        base = [r for r in rows if r[1:] == (0, 0, 0)][0][0]
        print(f"{wl} {name}: heuristic {base:.2f} us/launch; best:", ", ".join(f"{r[0]:.2f}us(st{r[1]} wpt{r[2]} ks{r[3]})" for r in rows[:6]), flush=True)
        out[name] = {"heuristic_us": base, "best": rows[:8]}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/qmv_sweep_{wl}.json", "w"), indent=1)

if __name__ == "__main__":
    main()
