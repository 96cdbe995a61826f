#!/usr/bin/env python
"""bench.py — world-step throughput of the B200-native forge2d step (BASELINE.json metric) on synthetic scenes.

Headline workload (config 5 of BASELINE.json): a batch of independent bench2d worlds (the exact scene of
packages/benchmark/bin/bench2d.dart: 40-row pyramid, 820 boxes, dt = 1/60 s, 4 sub-steps), every world pre-rolled
256 frames like the Dart harness's warm-up, sharded by world across the ranks with NO collective on the step path.
One "step" = one b2World_Step of every world of the batch = one launch of the one-CTA-per-world kernel.

  value     world-steps/s, whole job, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       the same through the C ABI with HOST buffers: per step an H2D of the per-world step inputs (gravity) from
            pinned memory, the step, and a D2H of every body's move event (transform) into pinned host memory
  roofline  algorithmic bytes of SURVEY §8(d) per world-step x worlds per launch / average launch duration
  cpu_baseline / --impl reference: the reference's own C (oracle/_ref, unmodified Box2D v3.1.1 as vendored by
            forge2d) stepping a bounded sample of the same worlds on all host cores, one single-worker world per thread

Extra keys report the single-world configurations (bench2d ms/frame through b2World_Step, large_pyramid and
many_pyramids body-steps/s) with the reference timed beside them.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from forge2d_b200 import _abi as A  # noqa: E402
from forge2d_b200 import scenes  # noqa: E402

DT, SUB = scenes.TIME_STEP, scenes.SUB_STEPS
PREROLL = 256  # bench2d.dart:53-57 warm-up frames
DECORRELATION_STEPS = 60  # untimed steps after the worlds were translated apart: long enough for them to run out of phase
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbox2d_ref.so")


# ------------------------------------------------------------------------------------------------ helpers
_emit = print  # replaced in main(): writes the JSON line to the real stdout


def shard(total, rank, world_size):
    """Contiguous block of worlds for `rank` (world w -> GPU w*G/total)."""
    lo = total * rank // world_size
    hi = total * (rank + 1) // world_size
    return lo, hi


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def algorithmic_bytes(counts):
    """compulsory bytes of one world-step (each input read once, each output written once)."""
    nb, ns, nc, nt, nj, moved = (counts[k] for k in ("awake_bodies", "awake_shapes", "awake_contacts", "touching",
                                                      "joints", "moved"))
    v_tree = 2 * ns + moved * max(1.0, math.log2(max(ns, 2)))
    return 176.0 * nb + 104.0 * ns + 360.0 * nc + 156.0 * nt + 140.0 * nj + 40.0 * v_tree


def world_counts(lib, world):
    """Per-step work counters of one world, read through the introspection ABI."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    snap = H.snapshot(lib, world, trees=False)
    bodies, contacts = snap["bodies"], snap["contacts"]
    awake_ids = set(snap["awake_order"].tolist())
    awake_shapes = int(sum(1 for b in snap["shapes"]["bodyId"] if int(b) in awake_ids))
    touching = int(snap["color_contact_counts"].sum())
    return {"awake_bodies": int(len(snap["awake_order"])), "awake_shapes": awake_shapes,
            "awake_contacts": touching + int(len(snap["awake_contacts"])), "touching": touching,
            "joints": int(snap["color_joint_counts"].sum()), "moved": int(len(snap["move_array"])),
            "bodies": int(len(bodies)), "contacts": int(len(contacts))}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_reference():
    if not os.path.exists(REF_SO):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        else:
            raise RuntimeError("oracle/_ref/libbox2d_ref.so missing (build it in the container with /root/reference)")
    lib = A.Library(REF_SO, "reference")
    lib.dll.tap_step_worlds.restype = C.c_double
    lib.dll.tap_step_worlds.argtypes = [C.POINTER(A.WorldId), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    lib.dll.tap_pool_create.restype = C.c_void_p
    lib.dll.tap_pool_create.argtypes = [C.c_int]
    lib.dll.tap_pool_destroy.argtypes = [C.c_void_p]
    lib.dll.tap_create_world_mt.restype = A.WorldId
    lib.dll.tap_create_world_mt.argtypes = [C.POINTER(A.WorldDef), C.c_void_p, C.c_int]
    lib.dll.tap_step_mt.argtypes = [A.WorldId, C.c_float, C.c_int, C.c_void_p]
    return lib


def world_x_offset(k):
    """Worlds of the batch are decorrelated as SURVEY §8(d) C5 proposes: world k is the bench2d scene translated by
    k * 2^-10 along x - the same physics, different floating-point rounding, so the replicas drift apart (the settling
    pile is chaotic) instead of marching through the phases of the step in lock-step, which no real batch does."""
    if os.environ.get("F2D_BENCH_IDENTICAL_WORLDS"):
        return 0.0
    return k * 2.0 ** -10


# ------------------------------------------------------------------------------------------------ reference arm
def reference_batch(ref, sample_worlds, steps, warmup, threads, min_seconds=0.0):
    """`sample_worlds` bench2d worlds, built at the x offsets of the batch's worlds (world_x_offset), at frame PREROLL, stepped on
    `threads` host threads, one world per thread at a time (1 Box2D worker each: the best case for the CPU, BASELINE.md
    §3.3). One "step" = every sampled world advanced `frames_per_step` frames, chosen so that the timed region of
    `steps` steps lasts at least `min_seconds`. Returns (world-steps/s, seconds, frames_per_step)."""
    made = [scenes.bench2d(ref, x_offset=world_x_offset(k)) for k in range(sample_worlds)]
    ids = (A.WorldId * sample_worlds)(*[s.world for s in made])
    ref.dll.tap_step_worlds(ids, sample_worlds, DT, SUB, PREROLL, threads)
    calibration = ref.dll.tap_step_worlds(ids, sample_worlds, DT, SUB, max(2, warmup), threads) / max(2, warmup)
    frames_per_step = max(1, int(math.ceil(min_seconds / max(1e-9, calibration * steps))))
    seconds = ref.dll.tap_step_worlds(ids, sample_worlds, DT, SUB, steps * frames_per_step, threads)
    for s in made:
        s.destroy()
    return sample_worlds * steps * frames_per_step / seconds, seconds, frames_per_step


def reference_single(ref, scene_name, kw, warm, timed, workers):
    """One world stepped by the reference with `workers` Box2D workers on a persistent pthread pool."""
    pool = ref.dll.tap_pool_create(workers) if workers > 1 else None
    create = (lambda wd: ref.dll.tap_create_world_mt(C.byref(wd), pool, workers)) if pool else None
    s = scenes.SCENES[scene_name](ref, create=create, **kw)
    step = (lambda: ref.dll.tap_step_mt(s.world, DT, SUB, pool)) if pool else s.step
    for _ in range(warm):
        step()
    samples, awake = [], 0
    for _ in range(timed):
        t0 = time.perf_counter()
        step()
        samples.append(time.perf_counter() - t0)
        awake += ref.b2World_GetAwakeBodyCount(s.world)
    s.destroy()
    if pool:
        ref.dll.tap_pool_destroy(pool)
    return samples, awake


def run_reference_arm(args, rank):
    if rank != 0:
        return
    ref = load_reference()
    cores = host_cores()
    sample = args.ref_worlds or cpu_sample_worlds(cores)
    value, seconds, frames = reference_batch(ref, sample, args.steps, args.warmup, cores, min_seconds=3.0)
    line = {
        "impl": "reference", "metric": "world_steps_per_s", "value": value, "unit": "world-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": cores, "kind": "reference",
                         "sample": "%d of the %d bench2d worlds (the first ones: built at their x offsets, frame %d), one step = %d frames of "
                                   "every sampled world, %d steps timed = %.1f s, one single-worker world per thread on %d "
                                   "threads" % (sample, args.worlds, PREROLL, frames, args.steps, seconds, cores)},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cores": cores, "cpu_model": cpu_model()},
    }
    _emit(json.dumps(line))


def cpu_sample_worlds(cores):
    """Worlds in the bounded CPU sample (the same for the cpu_baseline leg and the reference arm)."""
    return max(cores, min(16 * cores, 256))


def workload_config(args):
    return {"workload": "bench2d_batch: %d independent bench2d worlds (40-row pyramid, 820 boxes, dt=1/60, 4 sub-steps) "
                        "pre-rolled %d frames, decorrelated (world k translated by k * 2^-10 along x), sharded by world across "
                        "ranks" % (args.worlds, PREROLL),
            "worlds": args.worlds, "bodies_per_world": 821, "substeps": SUB, "parallelism": "world-sharded x%d" % args.gpus,
            "l2": "inputs larger than L2 (every world image is touched once per step; batch >> 126 MB)"}


# ------------------------------------------------------------------------------------------------ B200 arm
def single_world_numbers(lib, ref, name, kw, warm, timed, mode, cores):
    """Single-world configuration through b2World_Step (synchronous C-ABI call, as World.step makes it), with the
    reference (1 worker = what forge2d ships; all cores = pthread pool) timed on the same scene."""
    s = scenes.SCENES[name](lib, **kw)
    lib.f2dWorld_SetLaunchMode(s.world, mode)
    for _ in range(warm):
        s.step()
    samples, awake = [], 0
    for _ in range(timed):
        t0 = time.perf_counter()
        s.step()
        samples.append(time.perf_counter() - t0)
        awake += lib.b2World_GetAwakeBodyCount(s.world)
    counts = world_counts(lib, s.world)
    err = lib.f2dGetLastError().decode()
    s.destroy()
    out = {"frames": timed, "warmup_frames": warm, "launch_mode": {0: "one CTA", 1: "cooperative grid", -1: "auto"}[mode],
           "ms_per_frame_mean": 1e3 * sum(samples) / timed, "ms_per_frame_p5": 1e3 * sorted(samples)[int(0.05 * timed)],
           "ms_per_frame_p95": 1e3 * sorted(samples)[int(0.95 * timed)], "body_steps_per_s": awake / sum(samples),
           "bodies": counts["bodies"], "contacts": counts["contacts"], "error": err or None}
    if ref is not None:
        for label, workers in (("ref_1worker", 1), ("ref_allcores", cores)):
            rs, ra = reference_single(ref, name, kw, warm, timed, workers)
            out[label] = {"workers": workers, "ms_per_frame_mean": 1e3 * sum(rs) / timed, "body_steps_per_s": ra / sum(rs)}
    return out


def game_loop_numbers(lib, frames=256, warm=256):
    """bench2d as a game drives it: per frame one b2World_Step, every body's transform read (b2World_GetBodyEvents, the
    way forge2d's bodyMoveEvents does; plus one b2Body_GetPosition / GetLinearVelocity) and one b2Body_ApplyForceToCenter.
    Reports ms/frame beside the bare step and the bytes that crossed PCIe per frame."""
    out = {}
    for label, interact in (("bare_step", False), ("read_all_transforms_and_push_one_body", True)):
        s = scenes.bench2d(lib)
        lib.f2dWorld_SetLaunchMode(s.world, 0)
        for _ in range(warm):
            s.step()
        h0, d0 = C.c_ulonglong(), C.c_ulonglong()
        lib.f2dGetTransferBytes(C.byref(h0), C.byref(d0))
        checksum = 0.0
        t0 = time.perf_counter()
        for f in range(frames):
            if interact:
                lib.b2Body_ApplyForceToCenter(s.bodies[1 + (37 * f) % 800], A.Vec2(40.0, 15.0), True)
            s.step()
            if interact:
                ev = lib.b2World_GetBodyEvents(s.world)
                checksum += ev.moveEvents[ev.moveCount - 1].transform.p.y
                checksum += lib.b2Body_GetPosition(s.bodies[5]).y + lib.b2Body_GetLinearVelocity(s.bodies[7]).y
        seconds = time.perf_counter() - t0
        h1, d1 = C.c_ulonglong(), C.c_ulonglong()
        lib.f2dGetTransferBytes(C.byref(h1), C.byref(d1))
        s.destroy()
        out[label] = {"ms_per_frame": 1e3 * seconds / frames, "h2d_bytes_per_frame": (h1.value - h0.value) / frames,
                      "d2h_bytes_per_frame": (d1.value - d0.value) / frames, "checksum": checksum}
    return out


def run_b200_arm(args, rank, world_size, local_rank):
    import forge2d_b200
    lib = forge2d_b200.load_library()
    if not lib.f2dSetDevice(local_rank):
        raise RuntimeError("bench: CUDA device %d not usable - forge2d_b200 has no CPU fallback" % local_rank)
    dist = None
    if world_size > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lo, hi = shard(args.worlds, rank, world_size)
    mine = hi - lo

    # template world pre-rolled like bench2d.dart's warm-up, then replicated into the batch (HBM resident)
    template = scenes.bench2d(lib)
    for _ in range(PREROLL):
        template.step()
    counts0 = world_counts(lib, template.world)
    batch = lib.f2dBatch_Create(template.world, mine)
    if not batch:
        raise RuntimeError("bench: f2dBatch_Create failed: %s" % lib.f2dGetLastError().decode())
    world_bytes = lib.f2dBatch_GetWorldBytes(batch)
    if args.batch_threads and not lib.f2dBatch_SetLaunchConfig(batch, args.batch_threads, args.batch_blocks_per_sm):
        raise RuntimeError("bench: unknown batch launch config %dx%d" % (args.batch_threads, args.batch_blocks_per_sm))

    # decorrelate the worlds (world_x_offset; global world index, so every rank count gives the same batch) and let
    # them drift out of phase before anything is timed
    offsets = (A.Vec2 * mine)(*[A.Vec2(world_x_offset(lo + i), 0.0) for i in range(mine)])
    lib.f2dBatch_TranslateWorlds(batch, offsets, mine)
    gravity = lib.f2dHostAlloc(8 * mine)
    grav = (A.Vec2 * mine).from_address(gravity)
    for i in range(mine):
        grav[i] = A.Vec2(0.0, -10.0)
    lib.f2dBatch_StepN(batch, DT, SUB, args.warmup + DECORRELATION_STEPS)
    lib.f2dBatch_Synchronize(batch)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.f2dWorld_GetKernelLaunchCount()
    barrier()
    lib.f2dBatch_Synchronize(batch)
    lib.f2dBatch_EventRecord(batch, 0)
    lib.f2dBatch_StepN(batch, DT, SUB, args.steps)
    lib.f2dBatch_EventRecord(batch, 1)
    lib.f2dBatch_Synchronize(batch)
    barrier()
    ms_local = lib.f2dBatch_EventElapsedMs(batch, 0, 1)
    launches = lib.f2dWorld_GetKernelLaunchCount() - launches0
    ms = max_over_ranks(ms_local)
    clocks = sampler.stop() if rank == 0 else None
    errors = lib.f2dBatch_GetErrorFlags(batch)

    # counters after the timed region (average with the ones before for the roofline's per-step work)
    scratch = scenes.bench2d(lib, rows=1)
    lib.f2dBatch_DownloadWorld(batch, 0, scratch.world)
    counts1 = world_counts(lib, scratch.world)
    bytes_per_world_step = 0.5 * (algorithmic_bytes(counts0) + algorithmic_bytes(counts1))

    # ---- end to end through the C ABI with host buffers
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    nb = counts0["bodies"]
    ev_ptr, cnt_ptr = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
    lib.f2dBatch_SetGravity(batch, gravity, mine)
    lib.f2dBatch_Step(batch, DT, SUB)
    lib.f2dBatch_ReadBodyEvents(batch, nb, C.byref(ev_ptr), C.byref(cnt_ptr))  # allocates the pinned staging
    barrier()
    # (a) the three calls one after the other: upload, step, read back
    t0 = time.perf_counter()
    moved = 0
    for _ in range(e2e_steps):
        lib.f2dBatch_SetGravity(batch, gravity, mine)
        lib.f2dBatch_Step(batch, DT, SUB)
        moved = lib.f2dBatch_ReadBodyEvents(batch, nb, C.byref(ev_ptr), C.byref(cnt_ptr))
    e2e_sequential_seconds = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # (b) the same work per step through the fused call: world slices on separate streams, each slice's events cross
    # PCIe while later slices are stepped; every step still ends with all events of that step on the host
    lib.f2dBatch_StepAndReadBodyEvents(batch, DT, SUB, nb, C.byref(ev_ptr), C.byref(cnt_ptr))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lib.f2dBatch_SetGravity(batch, gravity, mine)
        moved = lib.f2dBatch_StepAndReadBodyEvents(batch, DT, SUB, nb, C.byref(ev_ptr), C.byref(cnt_ptr))
    e2e_sliced_seconds = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # (c) the pipelined call (the headline): per step an H2D of the inputs, the step, and the D2H of every body's transform
    # (16 bytes per body) into pinned host memory; call k returns the transforms of step k-1, whose copy ran while step k
    # was computed. The timed region ends with the flush, so it covers e2e_steps steps AND e2e_steps read-backs.
    rec_ptr, cnt2_ptr = C.c_void_p(), C.POINTER(C.c_int)()
    lib.f2dBatch_SetGravity(batch, gravity, mine)
    lib.f2dBatch_StepPipelined(batch, DT, SUB, nb, 1, C.byref(rec_ptr), C.byref(cnt2_ptr))
    lib.f2dBatch_FlushPipelined(batch, C.byref(rec_ptr), C.byref(cnt2_ptr))
    barrier()
    t0 = time.perf_counter()
    e2e_step_ms = []
    checksum = 0.0
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        lib.f2dBatch_SetGravity(batch, gravity, mine)
        moved = lib.f2dBatch_StepPipelined(batch, DT, SUB, nb, 1, C.byref(rec_ptr), C.byref(cnt2_ptr))
        if rec_ptr.value:
            checksum += C.cast(rec_ptr, C.POINTER(C.c_float))[1]  # the host really reads the result (y of the first body)
        e2e_step_ms.append(round(1e3 * (time.perf_counter() - t1), 2))
    moved = lib.f2dBatch_FlushPipelined(batch, C.byref(rec_ptr), C.byref(cnt2_ptr))
    checksum += C.cast(rec_ptr, C.POINTER(C.c_float))[1]
    e2e_seconds = max_over_ranks(time.perf_counter() - t0)
    growths = lib.f2dBatch_GetGrowthCount(batch)
    barrier()
    e2e_value = args.worlds * e2e_steps / e2e_seconds
    e2e_sequential_value = args.worlds * e2e_steps / e2e_sequential_seconds
    errors |= lib.f2dBatch_GetErrorFlags(batch)
    lib.f2dHostFree(gravity)
    lib.f2dBatch_Destroy(batch)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ms_per_step = ms / args.steps
    achieved = bytes_per_world_step * mine / (ms_per_step * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if prof.get("worlds_per_launch"):
            traffic = prof["dram_bytes_per_launch"] * mine / prof["worlds_per_launch"]
    except (OSError, ValueError, KeyError):
        pass

    line = {
        "metric": "world_steps_per_s", "value": args.worlds * args.steps / (ms * 1e-3), "unit": "world-steps/s",
        "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "world-steps/s", "h2d_bytes_per_step": 8 * mine,
                "d2h_bytes_per_step": mine * nb * 16 + 4 * (mine + 1), "steps": e2e_steps,
                "path": "f2dBatch_SetGravity (pinned H2D of per-world inputs) + f2dBatch_StepPipelined (step, then pinned D2H "
                        "of every body's transform, 16 B each; call k returns step k-1's transforms, whose copy overlapped "
                        "step k; the timed region ends with f2dBatch_FlushPipelined), per rank; transforms read per step: "
                        "%d; checksum %.6g" % (moved, checksum),
                "ms_per_step_each": e2e_step_ms,
                "full_events_sliced_value": args.worlds * e2e_steps / e2e_sliced_seconds,
                "full_events_sliced_path": "f2dBatch_SetGravity + f2dBatch_StepAndReadBodyEvents: every 40-byte b2BodyMoveEvent "
                                           "of the step on the host when the call returns (world slices on separate streams), "
                                           "%d B per step" % (mine * nb * C.sizeof(A.BodyMoveEvent) + 4 * mine),
                "sequential_calls_value": e2e_sequential_value,
                "sequential_calls_path": "f2dBatch_SetGravity + f2dBatch_Step + f2dBatch_ReadBodyEvents, nothing overlapped"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": ("stepWorldsCta<%d,%d> (whole world step, one CTA per world)" % (args.batch_threads, args.batch_blocks_per_sm))
                     if args.batch_threads else "stepWorldsGang<128,7> (whole world step, seven worlds per CTA, phase-aligned)",
                     "algorithmic_bytes_per_world_step": bytes_per_world_step, "worlds_per_launch": mine,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "note": "random 32-byte sector gathers: the narrowphase keeps DRAM busy 96 %% of its time at 3.9 TB/s of "
                             "sector traffic, the other phases are latency-bound (profiles/README.md); %.1f MB image per "
                             "world" % (world_bytes / 1e6)},
        "clocks": clocks,
        "batch": {"worlds_this_rank": mine, "bytes_per_world_image": int(world_bytes), "error_flags": int(errors),
                  "image_growths": int(growths),
                  "counts_before": counts0, "counts_after": counts1},
    }

    ref = None
    cores = host_cores()
    if world_size == 1 and not args.no_cpu_baseline:
        ref = load_reference()
        sample = args.ref_worlds or cpu_sample_worlds(cores)
        v, seconds, frames = reference_batch(ref, sample, args.steps, 2, cores, min_seconds=3.0)
        line["cpu_baseline"] = {"value": v, "unit": "world-steps/s", "cores": cores, "kind": "reference",
                                "sample": "%d of the %d bench2d worlds (the first ones: built at their x offsets, frame %d), %d steps of "
                                          "%d frames each = %.1f s, one single-worker world per thread on %d threads (%s)"
                                          % (sample, args.worlds, PREROLL, args.steps, frames, seconds, cores, cpu_model())}
        line["host"] = {"cores": cores, "cpu_model": cpu_model()}
    if world_size == 1 and not args.no_extras:
        extras = {}
        extras["bench2d"] = single_world_numbers(lib, ref, "bench2d", {}, 256, 256, 0, cores)
        # frames 150-400: the bottom row reaches the ground at frame ~147, so this window is the impact and the collapse
        # (hundreds of begin / end touch events per step, island merges, continuous collision), not the free fall
        extras["large_pyramid"] = single_world_numbers(lib, ref, "large_pyramid", {}, 150, 250, 1, cores)
        extras["large_pyramid"]["window"] = "frames 150-400 (impact and collapse)"
        extras["many_pyramids_awake"] = single_world_numbers(lib, ref, "many_pyramids", {}, 2, 24, 1, cores)
        extras["joint_grid"] = single_world_numbers(lib, ref, "joint_grid", {}, 8, 32, 1, cores)
        extras["bench2d_game_loop"] = game_loop_numbers(lib)
        line["single_world"] = extras
        # latency floor of a synchronous b2World_Step for small scenes (north_star's "graph-replay latency": the whole
        # step is ONE kernel launch, so this is launch + header read-back + synchronise): a world with no awake body,
        # and a 55-box pyramid
        line["small_scene_latency"] = {
            "launches_per_step": 1,
            "static_only_world_ms_per_frame": single_world_numbers(lib, None, "bench2d", {"rows": 0}, 64, 256, 0, cores)["ms_per_frame_mean"],
            "bench2d_10_rows": single_world_numbers(lib, ref, "bench2d", {"rows": 10}, 128, 256, 0, cores),
        }
    _emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=8192, help="total worlds of the batch (sharded across ranks)")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--batch-threads", type=int, default=0,
                    help="0: the default batch kernel (several worlds per block, phase-aligned); else threads per world of "
                         "the one-world-per-block kernel")
    ap.add_argument("--batch-blocks-per-sm", type=int, default=8, help="resident worlds per SM the kernel is built for")
    ap.add_argument("--ref-worlds", type=int, default=0, help="worlds in the CPU sample (0 = scaled to the host cores)")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-world configurations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line: the JSON the driver parses. Libraries write there too (NCCL prints its version
    # banner to fd 1 at init), so fd 1 is pointed at stderr for the run and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global _emit
    _emit = lambda text: os.write(json_fd, (text + "\n").encode())
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_b200_arm(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
