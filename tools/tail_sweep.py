"""Per-launch fixed cost of the megakernel at 1/8 frame: config 3 as ONE launch vs as the 8 balanced row tiles of the N = 8
bench run, for a sweep of the tile-scheduling knobs (each setting in its own process: the plugin reads them once).
usage: tail_sweep.py NAME=v1,v2 [NAME2=w1,w2 ...]   (the cross product is run)"""
import itertools, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import importlib, json, os, sys
sys.path.insert(0, %(root)r)
rtb = importlib.import_module("raytracing-in-one-weekend_b200")
W, H, spp = 1920, 1080, 256
scene = rtb.host.make_scene("final", max_bvh_depth=16)
ctx = rtb.plugin.Context(0); ctx.upload(scene)
b = rtb.plugin.HostBuffers(W, H); ctx.register_host_buffers(b)
def ms(r0, r1, reps=2):
    p = rtb.host.make_params(scene, W, H, spp, int(os.environ.get("SWEEP_TRACE_DEPTH", "50")), aperture=0.1, row_begin=r0, row_end=r1)
    best = 1e9
    for _ in range(reps):
        ctx.sample_batch(p, b); best = min(best, ctx.last_kernel_ms())
    return best
bounds = [0, 126, 230, 324, 417, 507, 609, 749, 1080]      # row tiles of the N = 8 run (profiles/r2_bench_n8.json)
full = ms(0, H)
tiles = [ms(bounds[i], bounds[i + 1], 3) for i in range(8)]
print("RESULT " + json.dumps({"tag": %(tag)r, "full": round(full, 2), "tiles_sum": round(sum(tiles), 2), "tiles_max": round(max(tiles), 2),
                              "tiles": [round(x, 2) for x in tiles]}))
'''


def main():
    axes = [a.split("=", 1) for a in sys.argv[1:]]
    names = [a[0] for a in axes]
    combos = list(itertools.product(*[a[1].split(",") for a in axes])) if axes else [()]
    for combo in combos:
        env = dict(os.environ)
        env.update(dict(zip(names, combo)))
        tag = " ".join(f"{n}={v}" for n, v in zip(names, combo)) or "default"
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "tag": tag}], env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
        print(line[0][7:] if line else f"{tag} FAILED {r.stdout[-300:]} {r.stderr[-800:]}", flush=True)


if __name__ == "__main__":
    main()
