"""Quick GPU timing probe (development aid): single-world ms/frame per launch mode and batch-step time vs batch size."""
import ctypes as C
import sys
import time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import forge2d_b200
from forge2d_b200 import scenes
from forge2d_b200 import _abi as A

lib = forge2d_b200.load_library()
assert lib.f2dHasDevice()
what = sys.argv[1:] or ["single", "batch"]
PROF_NAMES = ["begin", "pairQuery", "pairCreate", "treeRebuild", "narrow", "statePass", "solveSetup", "prepare", "integrateVel",
              "warmStart", "solve", "integratePos", "relax", "restitution", "store", "finalizeBodies", "hitEvents", "enlarge",
              "bullets", "sleep", "end", "splitJoin", "splitApply", "treeBeside", "tree.collect", "tree.positions",
              "tree.levels", "tree.tail", "tree.refit"]


def single(name, kw, mode, warm, timed):
    s = scenes.SCENES[name](lib, **kw)
    lib.f2dWorld_SetLaunchMode(s.world, mode)
    lib.f2dWorld_EnablePhaseTiming(s.world, True)
    t0 = time.perf_counter()
    s.step()
    first = time.perf_counter() - t0
    for _ in range(warm):
        s.step()
    acc = [0.0] * 5
    t0 = time.perf_counter()
    for _ in range(timed):
        s.step()
        t = (C.c_float * 5)()
        lib.f2dWorld_GetLastStepTimes(s.world, t)
        for i in range(5):
            acc[i] += t[i]
    wall = (time.perf_counter() - t0) / timed * 1e3
    print("%-16s %-22s mode %d: first %.2f ms, wall %.3f ms/frame, device phases [pairs %.3f collide %.3f solve %.3f finalize %.3f] total %.3f ms; err=%r"
          % (name, kw, mode, first * 1e3, wall, *[a / timed for a in acc], lib.f2dGetLastError()), flush=True)
    s.destroy()


if "single" in what:
    single("bench2d", {}, 0, 256, 64)
    single("bench2d", {}, 1, 256, 64)
    single("bench2d", dict(rows=10), 0, 64, 64)
    single("large_pyramid", {}, 1, 16, 16)
    single("large_pyramid", {}, 0, 4, 8)
    single("many_pyramids", {}, 1, 2, 8)
    single("joint_grid", {}, 1, 4, 8)

if "batch" in what:
    t = scenes.bench2d(lib)
    for _ in range(256):
        t.step()
    for count in (148, 296, 592, 1184, 2368):
        b = lib.f2dBatch_Create(t.world, count)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 2)
        lib.f2dBatch_Synchronize(b)
        lib.f2dBatch_EventRecord(b, 0)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 4)
        lib.f2dBatch_EventRecord(b, 1)
        lib.f2dBatch_Synchronize(b)
        ms = lib.f2dBatch_EventElapsedMs(b, 0, 1) / 4
        print("batch %5d worlds: %.3f ms/step -> %.0f world-steps/s (image %.2f MB) err=%x" % (
            count, ms, count / ms * 1e3, lib.f2dBatch_GetWorldBytes(b) / 1e6, lib.f2dBatch_GetErrorFlags(b)), flush=True)
        lib.f2dBatch_Destroy(b)

_PROF_NAMES_MOVED = ["begin", "pairQuery", "pairCreate", "treeRebuild", "narrow", "statePass", "solveSetup", "prepare", "integrateVel",
              "warmStart", "solve", "integratePos", "relax", "restitution", "store", "finalizeBodies", "hitEvents", "enlarge",
              "bullets", "sleep", "end", "splitJoin", "splitApply", "treeBeside", "tree.collect", "tree.positions",
              "tree.levels", "tree.tail", "tree.refit"]


def profile(name, kw, mode, warm, timed):
    s = scenes.SCENES[name](lib, **kw)
    lib.f2dWorld_SetLaunchMode(s.world, mode)
    for _ in range(warm):
        s.step()
    lib.f2dWorld_EnableProfile(s.world, True)
    t0 = time.perf_counter()
    for _ in range(timed):
        s.step()
    wall = (time.perf_counter() - t0) / timed * 1e3
    out = (C.c_ulonglong * 32)()
    lib.f2dWorld_ReadProfile(s.world, out, 32)
    total = sum(out[:23]) / timed / 1e3
    if out[31]:
        print("   query clock: max-accumulated %d cycles, mean per query %.0f cycles over %d queries (%.1f per frame)" % (out[29], out[30] / out[31], out[31], out[31] / timed))
    print("%s %s mode %d: wall %.3f ms/frame, in-kernel %.1f us: " % (name, kw, mode, wall, total) +
          " ".join("%s=%.1f" % (n, out[i] / timed / 1e3) for i, n in enumerate(PROF_NAMES) if out[i]), flush=True)
    s.destroy()


if "configs" in what:
    t = scenes.bench2d(lib)
    for _ in range(256):
        t.step()
    for threads, bps in ((256, 4), (128, 8), (64, 16)):
        count = 148 * bps * (2 if threads <= 64 else 4)
        b = lib.f2dBatch_Create(t.world, count)
        assert lib.f2dBatch_SetLaunchConfig(b, threads, bps)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 2)
        lib.f2dBatch_Synchronize(b)
        lib.f2dBatch_EventRecord(b, 0)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 4)
        lib.f2dBatch_EventRecord(b, 1)
        lib.f2dBatch_Synchronize(b)
        ms = lib.f2dBatch_EventElapsedMs(b, 0, 1) / 4
        print("config %3d threads x %2d blocks/SM, %5d worlds: %.3f ms/step -> %.0f world-steps/s err=%x" % (
            threads, bps, count, ms, count / ms * 1e3, lib.f2dBatch_GetErrorFlags(b)), flush=True)
        lib.f2dBatch_Destroy(b)

if "batchprofile" in what:
    t = scenes.bench2d(lib)
    for _ in range(256):
        t.step()
    lib.f2dWorld_EnableProfile(t.world, True)
    for threads, bps, count in ((64, 16, 2368), (128, 8, 1184), (256, 4, 592)):
        b = lib.f2dBatch_Create(t.world, count)
        assert lib.f2dBatch_SetLaunchConfig(b, threads, bps)
        steps = 8
        lib.f2dBatch_EventRecord(b, 0)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, steps)
        lib.f2dBatch_EventRecord(b, 1)
        lib.f2dBatch_Synchronize(b)
        ms = lib.f2dBatch_EventElapsedMs(b, 0, 1) / steps
        scratch = scenes.bench2d(lib, rows=1)
        lib.f2dBatch_DownloadWorld(b, count // 2, scratch.world)
        out = (C.c_ulonglong * 32)()
        lib.f2dWorld_ReadProfile(scratch.world, out, 32)
        total = sum(out[:23]) / steps / 1e3
        print("batch %dx%d, %d worlds: %.3f ms/step (%.0f world-steps/s); world %d in-kernel %.1f us: " % (
            threads, bps, count, ms, count / ms * 1e3, count // 2, total) +
            " ".join("%s=%.1f" % (n, out[i] / steps / 1e3) for i, n in enumerate(PROF_NAMES) if out[i]), flush=True)
        lib.f2dBatch_Destroy(b)
        scratch.destroy()

if "gangprofile" in what:
    # in-kernel phase times of worlds of a decorrelated 8192-world batch: the default (gang) kernel and one world per block
    t = scenes.bench2d(lib)
    for _ in range(256):
        t.step()
    lib.f2dWorld_EnableProfile(t.world, True)
    for gang in (1, 0):
        count = 8192
        b = lib.f2dBatch_Create(t.world, count)
        if not gang:
            lib.f2dBatch_SetLaunchConfig(b, 128, 8)
        offsets = (A.Vec2 * count)(*[A.Vec2(k * 2.0 ** -10, 0.0) for k in range(count)])
        lib.f2dBatch_TranslateWorlds(b, offsets, count)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, 64)
        lib.f2dBatch_Synchronize(b)
        scratch = scenes.bench2d(lib, rows=1)
        before = {}
        for idx in (100, 4000, 8000):
            lib.f2dBatch_DownloadWorld(b, idx, scratch.world)
            out = (C.c_ulonglong * 32)()
            lib.f2dWorld_ReadProfile(scratch.world, out, 32)
            before[idx] = list(out)
        steps = 8
        lib.f2dBatch_EventRecord(b, 0)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, steps)
        lib.f2dBatch_EventRecord(b, 1)
        lib.f2dBatch_Synchronize(b)
        ms = lib.f2dBatch_EventElapsedMs(b, 0, 1) / steps
        for idx in (100, 4000, 8000):
            lib.f2dBatch_DownloadWorld(b, idx, scratch.world)
            out = (C.c_ulonglong * 32)()
            lib.f2dWorld_ReadProfile(scratch.world, out, 32)
            d = [(out[i] - before[idx][i]) / steps / 1e3 for i in range(len(PROF_NAMES))]
            print("%s, 8192 decorrelated worlds: %.3f ms/step; world %d in-kernel %.1f us: " % (
                "gang 128x7" if gang else "one world per block 128x8", ms, idx, sum(d[:23])) +
                " ".join("%s=%.1f" % (n, d[i]) for i, n in enumerate(PROF_NAMES) if d[i] > 0.05), flush=True)
        lib.f2dBatch_Destroy(b)
        scratch.destroy()

if "impact" in what:
    profile("large_pyramid", {}, 1, 150, 100)
    profile("large_pyramid", {}, 1, 260, 100)
    profile("bench2d", dict(rows=10), 0, 128, 128)

if "profile" in what:
    profile("bench2d", {}, 0, 256, 64)
    profile("bench2d", {}, 1, 256, 64)
    profile("large_pyramid", {}, 1, 32, 32)
    profile("large_pyramid", {}, 0, 32, 32)
    profile("many_pyramids", {}, 1, 2, 16)
    profile("joint_grid", {}, 1, 8, 32)

if "classes" in what:
    # narrowphase of mixed-shape worlds with / without the work list binned by pair class
    for name, kw, frames in (("falling_shapes", dict(count=600), 260), ("random_world", dict(seed=3, count=400), 200)):
        for binning in (True, False):
            s = scenes.SCENES[name](lib, **kw)
            lib.f2dWorld_SetLaunchMode(s.world, 0)
            lib.f2dWorld_EnablePairClassBinning(s.world, binning)
            for _ in range(frames):
                s.step()
            lib.f2dWorld_EnableProfile(s.world, True)
            timed = 64
            for _ in range(timed):
                s.step()
            out = (C.c_ulonglong * 32)()
            lib.f2dWorld_ReadProfile(s.world, out, 32)
            cnt = lib.b2World_GetCounters(s.world)
            print("%s %s binning %s: contacts %d, narrow %.1f us per frame, frame %.1f us" % (
                name, kw, binning, cnt.contactCount, out[4] / timed / 1e3, sum(out[:23]) / timed / 1e3), flush=True)
            s.destroy()

if "b2d" in what:
    for _ in range(3):
        profile("bench2d", {}, 0, 256, 64)
    profile("bench2d", dict(rows=10), 0, 64, 64)

if "occupancy" in what:
    # the same 128x8 kernel with 1, 2, 4, 8 worlds resident per SM: how much of a world's step time is contention
    t = scenes.bench2d(lib)
    for _ in range(256):
        t.step()
    lib.f2dWorld_EnableProfile(t.world, True)
    for threads, bps, count in ((128, 8, 148), (128, 8, 296), (128, 8, 592), (128, 8, 1184), (128, 8, 2368), (256, 4, 148), (128, 4, 592)):
        b = lib.f2dBatch_Create(t.world, count)
        assert lib.f2dBatch_SetLaunchConfig(b, threads, bps)
        steps = 8
        lib.f2dBatch_EventRecord(b, 0)
        lib.f2dBatch_StepN(b, scenes.TIME_STEP, scenes.SUB_STEPS, steps)
        lib.f2dBatch_EventRecord(b, 1)
        lib.f2dBatch_Synchronize(b)
        ms = lib.f2dBatch_EventElapsedMs(b, 0, 1) / steps
        scratch = scenes.bench2d(lib, rows=1)
        lib.f2dBatch_DownloadWorld(b, count // 2, scratch.world)
        out = (C.c_ulonglong * 32)()
        lib.f2dWorld_ReadProfile(scratch.world, out, 32)
        total = sum(out[:23]) / steps / 1e3
        print("occupancy %dx%d, %d worlds: %.3f ms/step (%.0f world-steps/s); world %d in-kernel %.1f us: " % (
            threads, bps, count, ms, count / ms * 1e3, count // 2, total) +
            " ".join("%s=%.1f" % (n, out[i] / steps / 1e3) for i, n in enumerate(PROF_NAMES) if out[i]), flush=True)
        lib.f2dBatch_Destroy(b)
        scratch.destroy()
