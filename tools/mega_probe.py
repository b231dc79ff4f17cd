"""Stage-by-stage probe of the persistent decode kernel on a BASELINE workload (prints after every stage so a hang is located):
    python tools/mega_probe.py <workload> [prefill] [steps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B

def say(*a):
    print(f"[{time.strftime('%H:%M:%S')}]", *a, flush=True)

workload = sys.argv[1]
prefill = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[workload][2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
mdir = bench.model_dir_for(workload)
say("checkpoint", mdir)
ctx = B.Context(0)
t0 = time.time()
eng = B.Engine(ctx, mdir, max_context_length=max(1024, prefill + 4 * steps + 64))
say(f"engine created in {time.time() - t0:.1f}s; persistent={eng.persistent_decode} reason='{eng.persistent_decode_reason}'")
try:
    eng.set_persistent_decode(True)
except Exception as ex:
    say("persistent kernel unavailable:", ex)
rng = np.random.default_rng(0)
prompt = rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32)
t0 = time.time()
first = eng.prefill(prompt)
say(f"prefill {prefill} tokens in {time.time() - t0:.2f}s -> {first}")
eng.snapshot()
for mode in (True, False, True):
    if mode and not eng.persistent_decode:
        try:
            eng.set_persistent_decode(True)
        except Exception as ex:
            say("cannot enable persistent mode:", ex)
            continue
    if not mode:
        eng.set_persistent_decode(False)
    eng.restore()
    t0 = time.time()
    s1 = eng.decode_timed(2)
    say(f"mode persistent={mode}: first 2 steps {s1 * 1e3:.3f} ms (wall {time.time() - t0:.2f}s)")
    eng.restore()
    n0 = eng.launch_count
    s = eng.decode_timed(steps)
    say(f"mode persistent={mode}: {steps} steps {s * 1e3 / steps:.4f} ms/step = {steps / s:.1f} tok/s, launches/step {(eng.launch_count - n0) / steps:.1f}")
    eng.restore()
    toks = [first]
    for _ in range(6):
        toks.append(eng.step_host(toks[-1]))
    say("tokens", toks)
# per-phase trace of the persistent kernel (CTA 0 and a far CTA): where a step's time goes
KIND = {1: "gemv", 3: "attn", 4: "act", 6: "dnupd", 7: "logits", 8: "finish"}
if eng.persistent_decode:
    for cta in (0, 97):
        eng.restore()
        eng.decode_timed(2)
        kinds, cyc = eng.decode_trace(cta)
        c = cyc.astype(np.float64)
        t_start = c[0, 0]
        total = (c[-1, 2] - t_start) / 1.965e3
        agg = {}
        for k, row in zip(kinds, c):
            a = agg.setdefault(KIND.get(int(k), str(k)), [0] + [0.0] * 7)
            a[0] += 1
            prev = row[0]
            for i in range(1, 8):
                if row[i] > 0:
                    a[i] += (row[i] - prev) / 1.965e3
                    prev = row[i]
        total = (c[-1, 6] - t_start) / 1.965e3
        say(f"trace cta {cta}: step {total:.1f} us over {len(kinds)} phases; columns = us between consecutive stamps [1..7] (7 = grid barrier)")
        for name, a in agg.items():
            say(f"   {name:7s} x{a[0]:4d}  " + "  ".join(f"{x / a[0]:6.2f}" for x in a[1:]) + f"   | sum {sum(a[1:]):8.1f} us")
eng.close(); ctx.close()
say("done")
