import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from uzu_b200 import binding as B
wl = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b-int4"
ctx = B.Context(); eng = B.Engine(ctx, bench.model_dir_for(wl), max_context_length=512)
for sel, name in {1: "mixer_in", 2: "mixer_out", 4: "up", 32: "up+gated", 8: "down", 16: "readout", 31: "all"}.items():
    r = []
    for nopdl in (0, 1):
        t, n = eng.time_linears(10, sel | (0x80000000 if nopdl else 0))
        r.append(t / n * 1e6)
    print(f"{wl} {name}: pdl {r[0]:.2f} us/launch, no-pdl {r[1]:.2f} us/launch", flush=True)
