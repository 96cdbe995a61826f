"""Profiling aid: steps a bench2d batch with the four coarse phases launched separately (F2D_PROFILE_PHASE_LAUNCHES=1),
to be run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,...`."""
import os
import sys
os.environ["F2D_PROFILE_PHASE_LAUNCHES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import forge2d_b200
from forge2d_b200 import scenes

lib = forge2d_b200.load_library()
worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t = scenes.bench2d(lib)
for _ in range(256):
    t.step()
b = lib.f2dBatch_Create(t.world, worlds)
lib.f2dBatch_SetLaunchConfig(b, 128, 8)  # the phase launches exist for the one-world-per-block kernel
for _ in range(steps):
    lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 1)
lib.f2dBatch_Synchronize(b)
print("errors %x" % lib.f2dBatch_GetErrorFlags(b))
