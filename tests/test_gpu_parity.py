"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libforge2d_b200.so), against the
compiled reference (oracle/_ref) on identical call sequences. Bar: bit-exact for every integer/index record (pairs,
manifold point counts, feature ids, colours, islands, events) AND bit-exact floats (tolerance 0) — the kernels are
built without FMA contraction to match the reference's SSE2 arithmetic."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from forge2d_b200 import _abi as A
from forge2d_b200 import scenes

pytestmark = pytest.mark.gpu


def _lockstep(ref, gpu, name, kw, frames, every, mode=-1, check_events=True):
    a = scenes.SCENES[name](ref, **kw)
    b = scenes.SCENES[name](gpu, **kw)
    gpu.f2dWorld_SetLaunchMode(b.world, mode)
    for f in range(frames):
        a.step()
        b.step()
        if check_events and (f % every == 0):
            ea, eb = H.events(ref, a.world), H.events(gpu, b.world)
            assert ea["begin"] == eb["begin"] and ea["end"] == eb["end"] and ea["hit"] == eb["hit"], "events, frame %d" % f
            assert (ea["moves"] == eb["moves"]).all(), "move events, frame %d" % f
        if f % every == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
            assert d == [], "%s frame %d: %s" % (name, f, d[:6])
            assert gpu.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_bench2d_small_every_frame(ref, gpu, mode):
    _lockstep(ref, gpu, "bench2d", dict(rows=12), 260, 1, mode)


def test_bench2d_full_300_frames(ref, gpu):
    _lockstep(ref, gpu, "bench2d", {}, 300, 10)


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_falling_shapes_all_manifold_functions(ref, gpu, mode):
    _lockstep(ref, gpu, "falling_shapes", dict(count=24), 240, 2, mode)


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_polygon_soup_eight_vertex_path(ref, gpu, mode):
    _lockstep(ref, gpu, "polygon_soup", dict(count=40), 220, 2, mode)


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_jointed_piles_split_with_joint_edges(ref, gpu, mode):
    _lockstep(ref, gpu, "jointed_piles", dict(chains=6), 300, 3, mode)


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_many_pyramids_sleep(ref, gpu, mode):
    _lockstep(ref, gpu, "many_pyramids", dict(grid=3, base=6), 120, 3, mode)


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_joint_grid_with_rain(ref, gpu, mode):
    _lockstep(ref, gpu, "joint_grid", dict(n=12, rain_every=4), 120, 3, mode)


@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_island_parallel_solve_path(ref, gpu, mode):
    """100 small piles (36 boxes each): one warp per island runs the whole sub-step loop; bit-identical every frame."""
    a = scenes.many_pyramids(ref, grid=10, base=8)
    b = scenes.many_pyramids(gpu, grid=10, base=8)
    gpu.f2dWorld_SetLaunchMode(b.world, mode)
    used = 0
    for f in range(45):
        a.step()
        b.step()
        info = (C.c_int * 8)()
        gpu.f2dWorld_GetStepInfo(b.world, info, 8)
        used += info[0]
        if f % 3 == 0:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
            assert d == [], "frame %d: %s" % (f, d[:6])
    assert used >= 20
    assert gpu.f2dGetLastError() == b""


def test_large_pyramid_grid_mode(ref, gpu):
    _lockstep(ref, gpu, "large_pyramid", {}, 40, 8, 1, check_events=False)


def test_wake_through_api_after_sleep(ref, gpu):
    a = scenes.many_pyramids(ref, grid=2, base=5)
    b = scenes.many_pyramids(gpu, grid=2, base=5)
    for f in range(90):
        a.step()
        b.step()
    assert gpu.b2World_GetAwakeBodyCount(b.world) == 0
    ref.b2Body_SetLinearVelocity(a.bodies[7], A.Vec2(3.0, 4.0))
    gpu.b2Body_SetLinearVelocity(b.bodies[7], A.Vec2(3.0, 4.0))
    for f in range(30):
        a.step()
        b.step()
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
        assert d == [], "frame %d after wake: %s" % (f, d[:6])


def test_large_pyramid_through_the_impact(ref, gpu):
    """5050 boxes (BASELINE config 2) in lock-step with the reference THROUGH the landing of the bottom row (frame ~147)
    and the collapse that follows: hundreds of begin / end touch events per step, island merges, the ordered contact
    state pass at scale, continuous collision. Events are compared every frame from 140 on, every record every 8th."""
    a = scenes.large_pyramid(ref)
    b = scenes.large_pyramid(gpu)
    contacts = {}
    for f in range(1, 301):
        a.step()
        b.step()
        if f >= 140:
            ea, eb = H.events(ref, a.world), H.events(gpu, b.world)
            assert ea["begin"] == eb["begin"] and ea["end"] == eb["end"] and ea["hit"] == eb["hit"], "frame %d: events" % f
            if f % 8 == 0 or f in (147, 148, 149, 150):
                sb = H.snapshot(gpu, b.world)
                d = H.diff(H.snapshot(ref, a.world), sb)
                assert d == [], "frame %d: %s" % (f, d[:6])
                contacts[f] = int(sb["color_contact_counts"].sum())
    assert gpu.f2dGetLastError() == b""
    # free fall: 9900 touching contacts (rows resting on each other); the landing adds the ground contacts and the
    # collapse keeps changing the set
    assert contacts[144] == 9900 and contacts[160] != 9900 and len(set(contacts.values())) > 10, contacts


def test_zero_dt_and_substep_variants(ref, gpu):
    a = scenes.bench2d(ref, rows=8)
    b = scenes.bench2d(gpu, rows=8)
    for f in range(60):
        dt, sub = ((0.0, 4) if f % 7 == 3 else (1.0 / 60.0, 1 + f % 5))
        a.step(dt, sub)
        b.step(dt, sub)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
        assert d == [], "frame %d: %s" % (f, d[:6])


def test_batch_worlds_match_reference_and_each_other(ref, gpu):
    """Config 5 in miniature: 296 replicated bench2d worlds stepped by the one-CTA-per-world kernel; sampled worlds are
    downloaded and compared with the reference; the move events of every world must be byte-identical (replicas)."""
    a = scenes.bench2d(ref, rows=16)
    b = scenes.bench2d(gpu, rows=16)
    count, frames = 296, 120
    batch = gpu.f2dBatch_Create(b.world, count)
    assert batch
    for f in range(frames):
        a.step()
    gpu.f2dBatch_StepN(batch, scenes.TIME_STEP, scenes.SUB_STEPS, frames)
    gpu.f2dBatch_Synchronize(batch)
    assert gpu.f2dBatch_GetErrorFlags(batch) == 0
    nb = 136
    events = (A.BodyMoveEvent * (count * nb))()
    counts = (C.c_int * count)()
    total = gpu.f2dBatch_GetBodyEvents(batch, events, nb, counts)
    assert total == count * nb and set(counts[:]) == {nb}
    raw = np.frombuffer(events, dtype=np.uint8).reshape(count, nb * C.sizeof(A.BodyMoveEvent))
    assert (raw == raw[0]).all()
    for index in (0, 147, count - 1):
        gpu.f2dBatch_DownloadWorld(batch, index, b.world)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
        assert d == [], "batch world %d: %s" % (index, d[:6])
    gpu.f2dBatch_Destroy(batch)


@pytest.mark.gpu
@pytest.mark.parametrize("config", [(128, 8), (256, 4), (64, 16), (512, 1)], ids=["128x8", "256x4", "64x16", "512x1"])
def test_batch_one_world_per_block_launch_configs(ref, gpu, config):
    """The one-world-per-block batch kernels (f2dBatch_SetLaunchConfig; the default is the gang kernel) and the
    one-world-per-SM launch of the single-world kernel: every world of the batch equals the reference."""
    a = scenes.bench2d(ref, rows=12)
    b = scenes.bench2d(gpu, rows=12)
    count, frames, nb = 60, 90, 79
    batch = gpu.f2dBatch_Create(b.world, count)
    assert batch
    assert gpu.f2dBatch_SetLaunchConfig(batch, config[0], config[1])
    for f in range(frames):
        a.step()
    gpu.f2dBatch_StepN(batch, scenes.TIME_STEP, scenes.SUB_STEPS, frames)
    gpu.f2dBatch_Synchronize(batch)
    assert gpu.f2dBatch_GetErrorFlags(batch) == 0
    events = (A.BodyMoveEvent * (count * nb))()
    counts = (C.c_int * count)()
    gpu.f2dBatch_GetBodyEvents(batch, events, nb, counts)
    raw = np.frombuffer(events, dtype=np.uint8).reshape(count, nb * C.sizeof(A.BodyMoveEvent))
    assert (raw == raw[0]).all()
    for index in (0, count - 1):
        gpu.f2dBatch_DownloadWorld(batch, index, b.world)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
        assert d == [], "config %s world %d: %s" % (config, index, d[:6])
    gpu.f2dBatch_Destroy(batch)


def test_batch_sliced_step_and_read_equals_step_then_read(ref, gpu):
    """f2dBatch_StepAndReadBodyEvents (world slices on separate streams, read-back overlapped with stepping) returns the
    bytes of f2dBatch_Step + f2dBatch_ReadBodyEvents, step after step, with per-world gravity uploaded before each step;
    a sampled world equals the reference stepped with the same gravity."""
    a = scenes.bench2d(ref, rows=12)
    b = scenes.bench2d(gpu, rows=12)
    count, frames, nb = 2400, 12, 79
    one = gpu.f2dBatch_Create(b.world, count)
    two = gpu.f2dBatch_Create(b.world, count)
    assert one and two
    gravity = (A.Vec2 * count)()
    size = nb * C.sizeof(A.BodyMoveEvent)
    for f in range(frames):
        g = A.Vec2(0.25 * (f % 3), -10.0 + 0.5 * (f % 2))
        for i in range(count):
            gravity[i] = g
        ref.b2World_SetGravity(a.world, g)
        a.step()
        for batch in (one, two):
            gpu.f2dBatch_SetGravity(batch, C.cast(gravity, C.c_void_p), count)
        ev1, cn1 = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
        gpu.f2dBatch_Step(one, scenes.TIME_STEP, scenes.SUB_STEPS)
        total1 = gpu.f2dBatch_ReadBodyEvents(one, nb, C.byref(ev1), C.byref(cn1))
        ev2, cn2 = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
        total2 = gpu.f2dBatch_StepAndReadBodyEvents(two, scenes.TIME_STEP, scenes.SUB_STEPS, nb, C.byref(ev2), C.byref(cn2))
        assert total1 == total2 == count * (nb - 1)  # every dynamic body moves; the static ground does not
        raw1 = np.ctypeslib.as_array(C.cast(ev1, C.POINTER(C.c_uint8)), shape=(count, size))
        raw2 = np.ctypeslib.as_array(C.cast(ev2, C.POINTER(C.c_uint8)), shape=(count, size))
        used = (nb - 1) * C.sizeof(A.BodyMoveEvent)
        assert (raw1[:, :used] == raw2[:, :used]).all(), "frame %d" % f
        assert cn1[0] == cn2[0] == cn2[count - 1] == nb - 1
    assert gpu.f2dBatch_GetErrorFlags(two) == 0
    for index in (0, 1234, count - 1):
        gpu.f2dBatch_DownloadWorld(two, index, b.world)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
        assert d == [], "batch world %d: %s" % (index, d[:6])
    gpu.f2dBatch_Destroy(one)
    gpu.f2dBatch_Destroy(two)


def test_async_step_then_synchronize(ref, gpu):
    a = scenes.bench2d(ref, rows=10)
    b = scenes.bench2d(gpu, rows=10)
    b.step()
    a.step()
    for f in range(50):
        a.step()
        gpu.f2dWorld_StepAsync(b.world, scenes.TIME_STEP, scenes.SUB_STEPS)
    gpu.f2dWorld_Synchronize(b.world)
    assert H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world)) == []


def test_game_loop_reads_and_forces_move_kilobytes_not_the_image(ref, gpu):
    """A frame as a game makes it: read every body's position (and some velocities), push one body, step. After a step
    the getters fetch the body arrays once (not the multi-megabyte image) and the force / impulse / velocity edits of
    awake bodies go up as the bytes they changed; results stay bit-identical to the reference."""
    a = scenes.bench2d(ref)
    b = scenes.bench2d(gpu)
    for s in (a, b):
        for _ in range(30):
            s.step()
    h2d0, d2h0 = C.c_ulonglong(), C.c_ulonglong()
    gpu.f2dGetTransferBytes(C.byref(h2d0), C.byref(d2h0))
    frames = 40
    for f in range(frames):
        pa = [(ref.b2Body_GetPosition(x).x, ref.b2Body_GetPosition(x).y) for x in a.bodies]
        pb = [(gpu.b2Body_GetPosition(x).x, gpu.b2Body_GetPosition(x).y) for x in b.bodies]
        assert pa == pb, "frame %d: positions" % f
        k = 1 + (37 * f) % 800
        va, vb = ref.b2Body_GetLinearVelocity(a.bodies[k]), gpu.b2Body_GetLinearVelocity(b.bodies[k])
        assert (va.x, va.y) == (vb.x, vb.y)
        for lib, s in ((ref, a), (gpu, b)):
            lib.b2Body_ApplyForceToCenter(s.bodies[k], A.Vec2(40.0, 15.0), True)
            if f % 4 == 1:
                lib.b2Body_ApplyLinearImpulseToCenter(s.bodies[k + 3], A.Vec2(-0.5, 1.0), True)
            if f % 4 == 2:
                lib.b2Body_ApplyTorque(s.bodies[k + 5], 3.0, True)
                lib.b2Body_ApplyAngularImpulse(s.bodies[k + 7], 0.25, True)
            if f % 4 == 3:
                lib.b2Body_SetLinearVelocity(s.bodies[k + 9], A.Vec2(0.5, 2.0))
                lib.b2Body_SetAngularVelocity(s.bodies[k + 11], -1.5)
                lib.b2Body_ApplyForce(s.bodies[k + 13], A.Vec2(5.0, 5.0), A.Vec2(0.1, 0.2), True)
                lib.b2Body_ApplyLinearImpulse(s.bodies[k + 15], A.Vec2(0.2, 0.1), A.Vec2(0.0, 0.3), True)
            s.step()
    h2d1, d2h1 = C.c_ulonglong(), C.c_ulonglong()
    gpu.f2dGetTransferBytes(C.byref(h2d1), C.byref(d2h1))
    up, down = (h2d1.value - h2d0.value) / frames, (d2h1.value - d2h0.value) / frames
    assert up < 4096, "host -> device bytes per frame: %d" % up            # a few dirty records, not a 4.4 MB image
    assert down < 400 * 1024, "device -> host bytes per frame: %d" % down  # body + sim + state arrays (~210 KB) + header
    d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, b.world))
    assert d == [], d[:6]
    assert gpu.f2dGetLastError() == b""


def test_profile_reports_the_phases_of_the_step(gpu):
    """b2World_GetProfile (box2d.h:169): the first call switches the in-kernel phase marks on, later calls return the
    milliseconds per phase averaged over the steps in between; the parts add up to the step."""
    s = scenes.bench2d(gpu, rows=20)
    for _ in range(20):
        s.step()
    first = gpu.b2World_GetProfile(s.world)
    assert first.step == 0.0
    for _ in range(30):
        s.step()
    p = gpu.b2World_GetProfile(s.world)
    assert p.step > 0.02 and p.pairs > 0.0 and p.collide > 0.0 and p.solveImpulses > 0.0 and p.relaxImpulses > 0.0 and p.transforms > 0.0
    assert abs(p.step - (p.pairs + p.collide + p.solve)) < 0.02 * p.step + 1e-3
    assert p.solve >= p.solveConstraints >= p.solveImpulses
    assert gpu.b2World_GetProfile(s.world).step == 0.0  # no step since the previous call
    s.destroy()
