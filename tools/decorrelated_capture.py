"""Profiling aid: a bench2d batch whose worlds have been decorrelated (world k translated by k * 2^-10 along x, bench.py world_x_offset) and
stepped until they run out of phase, for `ncu -k regex:stepWorldsCta -s <256 + warm> -c 1 --set full`."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import forge2d_b200
from forge2d_b200 import scenes, _abi as A

lib = forge2d_b200.load_library()
worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 64
jitter = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
t = scenes.bench2d(lib)
for _ in range(256):
    t.step()
b = lib.f2dBatch_Create(t.world, worlds)
if os.environ.get("F2D_ONE_WORLD_PER_BLOCK"):
    lib.f2dBatch_SetLaunchConfig(b, 128, 8)
if jitter:
    offsets = (A.Vec2 * worlds)(*[A.Vec2(i * 2.0 ** -10, 0.0) for i in range(worlds)])
    lib.f2dBatch_TranslateWorlds(b, offsets, worlds)
for _ in range(warm + 2):
    lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 1)
lib.f2dBatch_Synchronize(b)
lib.f2dBatch_EventRecord(b, 0)
lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 4)
lib.f2dBatch_EventRecord(b, 1)
lib.f2dBatch_Synchronize(b)
print("worlds %d jitter %s: %.3f ms/step after %d steps, errors %x" % (worlds, jitter, lib.f2dBatch_EventElapsedMs(b, 0, 1) / 4, warm + 2,
                                                                      lib.f2dBatch_GetErrorFlags(b)))
