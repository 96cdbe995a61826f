"""Development aid: one large seeded random world (hundreds of bodies) on the GPU in both launch modes against oracle/_ref."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import test_random_worlds as T  # noqa: E402
from forge2d_b200 import scenes  # noqa: E402

ref, gpu = H.load("reference"), H.load("product")
count, frames = int(sys.argv[1]), int(sys.argv[2])
for seed in (101, 102, 103):
    for mode in (0, 1):
        a = scenes.random_world(ref, seed=seed, count=count)
        b = scenes.random_world(gpu, seed=seed, count=count)
        gpu.f2dWorld_SetLaunchMode(b.world, mode)
        ok = True
        for f in range(frames):
            a.step()
            b.step()
            if T._events(ref, a.world) != T._events(gpu, b.world):
                print("seed %d mode %d frame %d: events differ" % (seed, mode, f)); ok = False; break
            if f % 10 == 0 or f == frames - 1:
                d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
                if d:
                    print("seed %d mode %d frame %d: %s" % (seed, mode, f, d[:4])); ok = False; break
        print("seed %d mode %d count %d: %s (%d contacts at the end) err=%r" % (
            seed, mode, count, "identical" if ok else "DIFFERS", len(H.snapshot(gpu, b.world)["contacts"]), gpu.f2dGetLastError()), flush=True)
        a.destroy(); b.destroy()
