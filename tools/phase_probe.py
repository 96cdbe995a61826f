"""In-kernel phase times (rank-0 globaltimer marks) of one world: python tools/phase_probe.py <scene> <mode> [k=v ...]"""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import forge2d_b200
from forge2d_b200 import scenes

NAMES = ["begin", "pairQuery", "pairCreate", "treeRebuild", "narrow", "statePass", "solveSetup", "prepare", "integrateVel", "warmStart",
         "solve", "integratePos", "relax", "restitution", "store", "finalizeBodies", "hitEvents", "enlarge", "bullets", "sleep", "end", "splitJoin", "splitApply"]
lib = forge2d_b200.load_library()
name, mode = sys.argv[1], int(sys.argv[2])
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:])}
s = scenes.SCENES[name](lib, **kw)
lib.f2dWorld_SetLaunchMode(s.world, mode)
for _ in range(200):
    s.step()
lib.f2dWorld_EnableProfile(s.world, True)
timed = 100
t0 = time.perf_counter()
for _ in range(timed):
    s.step()
wall = (time.perf_counter() - t0) / timed * 1e3
out = (C.c_ulonglong * 24)()
lib.f2dWorld_ReadProfile(s.world, out, 24)
print("%s %s mode %d: wall %.3f ms/frame, in-kernel %.1f us: " % (name, kw, mode, wall, sum(out[:23]) / timed / 1e3) +
      " ".join("%s=%.1f" % (n, out[i] / timed / 1e3) for i, n in enumerate(NAMES) if out[i]))
