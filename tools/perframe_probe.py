import ctypes as C, sys, time, os
sys.path.insert(0, '/root/repo')
import forge2d_b200
from forge2d_b200 import scenes
lib = forge2d_b200.load_library()
for name, warm, timed in (("joint_grid", 8, 32), ("many_pyramids", 2, 12)):
    s = scenes.SCENES[name](lib)
    lib.f2dWorld_SetLaunchMode(s.world, 1)
    for _ in range(warm):
        s.step()
    lib.f2dWorld_EnableProfile(s.world, True)
    prev = [0] * 32
    rows = []
    for f in range(timed):
        t0 = time.perf_counter()
        s.step()
        wall = (time.perf_counter() - t0) * 1e3
        out = (C.c_ulonglong * 32)()
        lib.f2dWorld_ReadProfile(s.world, out, 32)
        cur = list(out)
        k = sum(cur[:23]) - sum(prev[:23])
        prev = cur
        rows.append("%.3f/%.3f" % (wall, k / 1e6))
    print(name, "wall/in-kernel ms per frame:", " ".join(rows), flush=True)
    s.destroy()
