"""Steps one world N frames (development aid for ncu captures): python tools/single_step.py <scene> <frames> <mode> [kwargs]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import forge2d_b200
from forge2d_b200 import scenes
lib = forge2d_b200.load_library()
name, frames, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kw = eval(sys.argv[4]) if len(sys.argv) > 4 else {}
s = scenes.SCENES[name](lib, **kw)
lib.f2dWorld_SetLaunchMode(s.world, mode)
for _ in range(frames):
    s.step()
print("done", lib.f2dGetLastError())
