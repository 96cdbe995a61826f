"""Development aid: differential sweep of seeded random worlds / random API calls on the GPU against oracle/_ref
(beyond the seeds in tests/test_random_worlds.py): python tools/gpu_fuzz_sweep.py <first seed> <last seed>"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import test_random_worlds as T  # noqa: E402

ref, gpu = H.load("reference"), H.load("product")
first, last = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(first, last + 1):
    for mode in (0, 1):
        try:
            T._run(ref, gpu, seed, 200, 10, mode=mode)
            T._fuzz(ref, gpu, seed, frames=120, mode=mode)
        except AssertionError as e:
            bad.append((seed, mode))
            print("FAIL seed %d mode %d: %s" % (seed, mode, str(e)[:300]), flush=True)
print("swept seeds %d..%d in both launch modes; failures: %s" % (first, last, bad))
