"""Workloads for compute-sanitizer (racecheck / synccheck / memcheck): small versions of the launch modes.
usage: compute-sanitizer --tool <tool> python tools/sanitize_targets.py <cta|grid|batch|batch1>"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import forge2d_b200
from forge2d_b200 import scenes

lib = forge2d_b200.load_library()
what = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
if what == "cta":        # one block per world: pyramid through landing, collapse-free settling and island splits
    s = scenes.bench2d(lib, rows=10)
    lib.f2dWorld_SetLaunchMode(s.world, 0)
    for _ in range(frames):
        s.step()
    f = scenes.falling_shapes(lib, count=24)
    lib.f2dWorld_SetLaunchMode(f.world, 0)
    for _ in range(frames):
        f.step()
elif what == "grid":     # cooperative grid: small piles (island-parallel solve, sleeping) and a pyramid
    s = scenes.many_pyramids(lib, grid=3, base=6)
    lib.f2dWorld_SetLaunchMode(s.world, 1)
    for _ in range(frames):
        s.step()
    p = scenes.bench2d(lib, rows=8)
    lib.f2dWorld_SetLaunchMode(p.world, 1)
    for _ in range(frames // 2):
        p.step()
elif what in ("batch", "batch1"):   # several worlds per block (the default) / one world per block
    t = scenes.bench2d(lib, rows=8)
    for _ in range(40):
        t.step()
    b = lib.f2dBatch_Create(t.world, 40)
    if what == "batch1":
        lib.f2dBatch_SetLaunchConfig(b, 128, 8)
    lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, max(4, frames // 6))
    lib.f2dBatch_Synchronize(b)
    print("batch errors %x" % lib.f2dBatch_GetErrorFlags(b))
print("done", what, lib.f2dGetLastError())
