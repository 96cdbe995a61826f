"""Batches of DIFFERENT worlds (f2dBatch_CreateFromWorlds): every world of the batch against the reference stepped the
same number of frames, all records bit-identical; and the batch growth path (a world whose new contacts do not fit
its image stops, the batch grows every image on the device, the world repeats the step)."""
import ctypes as C

import pytest

import harness as H
from forge2d_b200 import _abi as A
from forge2d_b200 import scenes

pytestmark = pytest.mark.gpu


def _events_of(lib, world):
    ev = H.events(lib, world)
    return ev["begin"], ev["end"], ev["hit"], ev["sensor_begin"], ev["sensor_end"]


def test_batch_of_distinct_random_worlds_matches_reference(ref, gpu):
    """64 different seeded random worlds (every shape / body type, joints, sensors, bullets; tests/test_random_worlds.py)
    plus pyramids of different heights in ONE batch: after 150 frames every world is downloaded and compared with the
    reference; contact / hit / sensor events of the last step are compared through the downloaded images too."""
    gpu.f2dClearLastError()
    builders = [lambda lib, s=s: scenes.random_world(lib, seed=100 + s) for s in range(56)]
    builders += [lambda lib, r=r: scenes.bench2d(lib, rows=r) for r in (3, 5, 8, 11, 14, 17, 20, 24)]
    refs = [b(ref) for b in builders]
    ours = [b(gpu) for b in builders]
    ids = (A.WorldId * len(ours))(*[s.world for s in ours])
    batch = gpu.f2dBatch_CreateFromWorlds(ids, len(ours))
    assert batch, gpu.f2dGetLastError()
    frames = 150
    for f in range(frames):
        for s in refs:
            s.step()
    gpu.f2dBatch_StepN(batch, scenes.TIME_STEP, scenes.SUB_STEPS, frames)
    gpu.f2dBatch_Synchronize(batch)
    flags = (C.c_uint32 * len(ours))()
    bad = gpu.f2dBatch_GetWorldErrors(batch, flags, len(ours))
    assert bad == 0, [hex(f) for f in flags]
    scratch = scenes.bench2d(gpu, rows=1)
    for index, a in enumerate(refs):
        gpu.f2dBatch_DownloadWorld(batch, index, scratch.world)
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, scratch.world))
        assert d == [], "batch world %d: %s" % (index, d[:6])
        assert _events_of(ref, a.world) == _events_of(gpu, scratch.world), "batch world %d: events" % index
    gpu.f2dBatch_Destroy(batch)
    assert gpu.f2dGetLastError() == b""


def _crowd(lib, n, seed):
    """n circles created almost on top of each other: the first step creates ~n^2/2 contacts at once."""
    s = scenes.bench2d(lib, rows=1, ground_half_width=30.0)
    sd = lib.b2DefaultShapeDef()
    for k in range(n):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(-2.0 + 0.04 * k + 0.001 * seed, -20.0 + 0.03 * (k % 7))
        b = lib.b2CreateBody(s.world, C.byref(bd))
        c = A.Circle(A.Vec2(0.0, 0.0), 0.5)
        lib.b2CreateCircleShape(b, C.byref(sd), C.byref(c))
    return s


def test_batch_grows_when_a_world_needs_more_contact_room(ref, gpu):
    """Two crowded worlds (thousands of new contacts in their first step, far beyond the head-room of the common layout)
    among calm ones: nobody freezes, nobody falls behind, the batch reports that it grew."""
    gpu.f2dClearLastError()
    builders = [lambda lib: scenes.bench2d(lib, rows=6), lambda lib: _crowd(lib, 90, 1), lambda lib: scenes.bench2d(lib, rows=9),
                lambda lib: _crowd(lib, 70, 2), lambda lib: scenes.bench2d(lib, rows=4)]
    refs = [b(ref) for b in builders]
    ours = [b(gpu) for b in builders]
    ids = (A.WorldId * len(ours))(*[s.world for s in ours])
    batch = gpu.f2dBatch_CreateFromWorlds(ids, len(ours))
    assert batch, gpu.f2dGetLastError()
    scratch = scenes.bench2d(gpu, rows=1)
    done = 0
    for frames in (1, 2, 37):
        for f in range(frames):
            for s in refs:
                s.step()
        gpu.f2dBatch_StepN(batch, scenes.TIME_STEP, scenes.SUB_STEPS, frames)
        gpu.f2dBatch_Synchronize(batch)
        done += frames
        assert gpu.f2dBatch_GetErrorFlags(batch) == 0
        for index, a in enumerate(refs):
            gpu.f2dBatch_DownloadWorld(batch, index, scratch.world)
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, scratch.world))
            assert d == [], "after %d frames, batch world %d: %s" % (done, index, d[:6])
    assert gpu.f2dBatch_GetGrowthCount(batch) >= 1
    # the fused step + read-back call takes the same path
    ev, cn = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
    for s in refs:
        s.step()
    gpu.f2dBatch_StepAndReadBodyEvents(batch, scenes.TIME_STEP, scenes.SUB_STEPS, 128, C.byref(ev), C.byref(cn))
    for index, a in enumerate(refs):
        gpu.f2dBatch_DownloadWorld(batch, index, scratch.world)
        assert H.diff(H.snapshot(ref, a.world), H.snapshot(gpu, scratch.world)) == []
    gpu.f2dBatch_Destroy(batch)
    assert gpu.f2dGetLastError() == b""


def test_pipelined_step_returns_the_previous_steps_results(ref, gpu):
    """f2dBatch_StepPipelined: call k returns what f2dBatch_Step + f2dBatch_ReadBodyEvents returned after step k-1, in
    both formats (40-byte events, 16-byte transforms); the flush returns the last step's."""
    import numpy as np
    b = scenes.bench2d(gpu, rows=12)
    count, frames, nb = 300, 10, 79
    one = gpu.f2dBatch_Create(b.world, count)
    two = gpu.f2dBatch_Create(b.world, count)
    three = gpu.f2dBatch_Create(b.world, count)
    offsets = (A.Vec2 * count)(*[A.Vec2(k * 2.0 ** -10, 0.0) for k in range(count)])
    for batch in (one, two, three):
        gpu.f2dBatch_TranslateWorlds(batch, offsets, count)
    size = nb * C.sizeof(A.BodyMoveEvent)
    used = (nb - 1) * C.sizeof(A.BodyMoveEvent)
    previous = None
    for f in range(frames + 1):
        rec, cnt = C.c_void_p(), C.POINTER(C.c_int)()
        rec3, cnt3 = C.c_void_p(), C.POINTER(C.c_int)()
        if f < frames:
            got = gpu.f2dBatch_StepPipelined(two, scenes.TIME_STEP, scenes.SUB_STEPS, nb, 0, C.byref(rec), C.byref(cnt))
            got3 = gpu.f2dBatch_StepPipelined(three, scenes.TIME_STEP, scenes.SUB_STEPS, nb, 1, C.byref(rec3), C.byref(cnt3))
        else:
            got = gpu.f2dBatch_FlushPipelined(two, C.byref(rec), C.byref(cnt))
            got3 = gpu.f2dBatch_FlushPipelined(three, C.byref(rec3), C.byref(cnt3))
        if f == 0:
            assert got == 0 and got3 == 0 and not rec.value
        else:
            assert got == got3 == count * (nb - 1)
            raw = np.ctypeslib.as_array(C.cast(rec, C.POINTER(C.c_uint8)), shape=(count, size))[:, :used].copy()
            assert (raw == previous).all(), "frame %d" % f
            # transforms: the first 16 bytes of every 40-byte event
            xf = np.ctypeslib.as_array(C.cast(rec3, C.POINTER(C.c_uint8)), shape=(count, nb * 16))[:, :(nb - 1) * 16]
            want = previous.reshape(count, nb - 1, C.sizeof(A.BodyMoveEvent))[:, :, :16].reshape(count, (nb - 1) * 16)
            assert (xf == want).all(), "frame %d: transforms" % f
            assert cnt[0] == cnt3[count - 1] == nb - 1
        if f < frames:
            ev1, cn1 = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
            gpu.f2dBatch_Step(one, scenes.TIME_STEP, scenes.SUB_STEPS)
            gpu.f2dBatch_ReadBodyEvents(one, nb, C.byref(ev1), C.byref(cn1))
            previous = np.ctypeslib.as_array(C.cast(ev1, C.POINTER(C.c_uint8)), shape=(count, size))[:, :used].copy()
    # translated replicas really are different worlds
    assert not (previous[0] == previous[count - 1]).all()
    for batch in (one, two, three):
        assert gpu.f2dBatch_GetErrorFlags(batch) == 0
        gpu.f2dBatch_Destroy(batch)
