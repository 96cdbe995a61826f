"""Differential test on seeded random scenes (forge2d_b200/scenes.py random_world): every body and shape type with random
materials, filters, sensors, bullets, multi-shape bodies and joints. The scene is built by the same call
sequence against the compiled reference and against our library and stepped side by side; every internal record must
stay bit-identical (pairs, manifolds, colours, islands, trees, float state) and so must the event streams.
CPU: host emulation of the step templates; `-m gpu`: the CUDA product in both launch modes."""
import pytest

import harness as H
from forge2d_b200 import scenes

SEEDS = [1, 2, 3, 4, 5, 6, 7, 8]


def _events(lib, world):
    c = lib.b2World_GetContactEvents(world)
    s = lib.b2World_GetSensorEvents(world)
    begin = [(c.beginEvents[i].shapeIdA.index1, c.beginEvents[i].shapeIdB.index1) for i in range(c.beginCount)]
    end = [(c.endEvents[i].shapeIdA.index1, c.endEvents[i].shapeIdB.index1) for i in range(c.endCount)]
    hit = [(c.hitEvents[i].shapeIdA.index1, c.hitEvents[i].shapeIdB.index1, c.hitEvents[i].point.x, c.hitEvents[i].point.y,
            c.hitEvents[i].approachSpeed) for i in range(c.hitCount)]
    sb = [(s.beginEvents[i].sensorShapeId.index1, s.beginEvents[i].visitorShapeId.index1) for i in range(s.beginCount)]
    se = [(s.endEvents[i].sensorShapeId.index1, s.endEvents[i].visitorShapeId.index1) for i in range(s.endCount)]
    return begin, end, hit, sb, se


def _run(ref, lib, seed, frames, every, mode=None):
    lib.f2dClearLastError()
    a = scenes.random_world(ref, seed=seed)
    b = scenes.random_world(lib, seed=seed)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    totals = [0, 0, 0, 0, 0]
    for f in range(frames):
        a.step()
        b.step()
        ea, eb = _events(ref, a.world), _events(lib, b.world)
        assert ea == eb, "seed %d frame %d: event streams differ" % (seed, f)
        totals = [t + len(x) for t, x in zip(totals, ea)]
        if f % every == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
            assert d == [], "seed %d frame %d: %s" % (seed, f, d[:6])
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()
    return totals


@pytest.mark.parametrize("seed", SEEDS)
def test_random_world_matches_reference_emu(ref, emu, seed):
    totals = _run(ref, emu, seed, 240, 6)
    assert totals[0] > 20      # contacts began


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
@pytest.mark.parametrize("seed", SEEDS[:5])
def test_random_world_matches_reference_gpu(ref, gpu, seed, mode):
    _run(ref, gpu, seed, 240, 12, mode=mode)


def _crowd(lib, n=90):
    """n circles created almost on top of each other: the first step creates ~n^2/2 contacts at once, far more than the
    head-room any image is laid out with (the reference's arrays simply grow)."""
    import ctypes as C
    from forge2d_b200 import _abi as A
    s = scenes.bench2d(lib, rows=1, ground_half_width=30.0)
    sd = lib.b2DefaultShapeDef()
    for k in range(n):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(-2.0 + 0.04 * k, -20.0 + 0.03 * (k % 7))
        b = lib.b2CreateBody(s.world, C.byref(bd))
        c = A.Circle(A.Vec2(0.0, 0.0), 0.5)
        lib.b2CreateCircleShape(b, C.byref(sd), C.byref(c))
    return s


def _crowd_session(ref, lib, mode=None):
    lib.f2dClearLastError()
    a, b = _crowd(ref), _crowd(lib)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    most = 0
    for f in range(40):
        a.step()
        b.step()
        if f < 3 or f % 8 == 0:
            sb = H.snapshot(lib, b.world)
            d = H.diff(H.snapshot(ref, a.world), sb)
            assert d == [], "frame %d: %s" % (f, d[:6])
            most = max(most, len(sb["contacts"]))
    assert most > 1500      # 90 bodies moved: the image had room for 4 * 91 + 64 new contacts
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_step_grows_the_image_when_new_contacts_do_not_fit_emu(ref, emu):
    _crowd_session(ref, emu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_step_grows_the_image_when_new_contacts_do_not_fit_gpu(ref, gpu, mode):
    _crowd_session(ref, gpu, mode=mode)
