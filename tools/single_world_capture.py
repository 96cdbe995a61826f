"""Profiling aid: one bench2d world stepped to its steady state in the one-block kernel, for
`ncu -k regex:stepWorldsCta -s 300 -c 1 --set full` (launch 300 = frame 300)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import forge2d_b200
from forge2d_b200 import scenes

lib = forge2d_b200.load_library()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 320
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 40
s = scenes.bench2d(lib, rows=rows)
lib.f2dWorld_SetLaunchMode(s.world, 0)
for _ in range(frames):
    s.step()
print("stepped", frames, "frames, error", lib.f2dGetLastError())
