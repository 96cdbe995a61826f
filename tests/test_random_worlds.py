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


def _register_callbacks(lib, world, log):
    """Deterministic b2CustomFilterFcn / b2PreSolveFcn (verdicts are pure functions of the arguments) that log what they see."""
    import ctypes as C
    from forge2d_b200 import _abi as A

    def custom_filter(a, b, ctx):
        log.append(("filter", a.index1, b.index1))
        return (a.index1 * 3 + b.index1) % 7 != 0

    def pre_solve(a, b, manifold, ctx):
        m = manifold.contents
        log.append(("presolve", a.index1, b.index1, m.pointCount, m.normal.x, m.normal.y, m.points[0].separation, m.points[0].id))
        return (a.index1 + 2 * b.index1) % 5 != 0

    f1 = C.CFUNCTYPE(C.c_bool, A.ShapeId, A.ShapeId, C.c_void_p)(custom_filter)
    f2 = C.CFUNCTYPE(C.c_bool, A.ShapeId, A.ShapeId, C.POINTER(A.Manifold), C.c_void_p)(pre_solve)
    lib.b2World_SetCustomFilterCallback(world, C.cast(f1, C.c_void_p), None)
    lib.b2World_SetPreSolveCallback(world, C.cast(f2, C.c_void_p), None)
    return f1, f2


def _run(ref, lib, seed, frames, every, mode=None, callbacks=False):
    lib.f2dClearLastError()
    a = scenes.random_world(ref, seed=seed, pre_solve_events=callbacks)
    b = scenes.random_world(lib, seed=seed, pre_solve_events=callbacks)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    logs, keep = ([], []), []
    if callbacks:
        keep = [_register_callbacks(ref, a.world, logs[0]), _register_callbacks(lib, b.world, logs[1])]
    totals = [0, 0, 0, 0, 0]
    for f in range(frames):
        a.step()
        b.step()
        ea, eb = _events(ref, a.world), _events(lib, b.world)
        assert ea == eb, "seed %d frame %d: event streams differ" % (seed, f)
        assert logs[0] == logs[1], "seed %d frame %d: callback sequences differ" % (seed, f)
        totals = [t + len(x) for t, x in zip(totals, ea)]
        if f % every == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
            assert d == [], "seed %d frame %d: %s" % (seed, f, d[:6])
    assert lib.f2dGetLastError() == b""
    if callbacks:
        assert sum(1 for e in logs[0] if e[0] == "presolve") > 50 and sum(1 for e in logs[0] if e[0] == "filter") > 50
    a.destroy()
    b.destroy()
    del keep
    return totals


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_world_with_step_callbacks_matches_reference_emu(ref, emu, seed):
    _run(ref, emu, seed, 200, 8, callbacks=True)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
@pytest.mark.parametrize("seed", [11, 12])
def test_random_world_with_step_callbacks_matches_reference_gpu(ref, gpu, seed, mode):
    _run(ref, gpu, seed, 200, 16, mode=mode, callbacks=True)


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


# ---- random API calls between steps --------------------------------------------------------------------------------
def _mutate(lib, scene, rng, alive, joints_alive):
    """One random mutator call (same choice for every library: the choice depends only on the seeded stream)."""
    import ctypes as C
    import math
    from forge2d_b200 import _abi as A
    r = rng.next
    pick = [i for i in range(1, len(scene.bodies)) if alive[i]]
    if not pick:
        return "none"
    i = pick[int(r() * len(pick))]
    body = scene.bodies[i]
    op = int(r() * 16)
    if op == 0:
        lib.b2Body_ApplyLinearImpulseToCenter(body, A.Vec2(H_f32(6.0 * r() - 3.0), H_f32(8.0 * r())), True)
    elif op == 1:
        a = 6.28318 * r()
        lib.b2Body_SetTransform(body, A.Vec2(H_f32(-7.0 + 14.0 * r()), H_f32(2.0 + 20.0 * r())), A.Rot(H_f32(math.cos(a)), H_f32(math.sin(a))))
    elif op == 2:
        lib.b2Body_SetLinearVelocity(body, A.Vec2(H_f32(10.0 * r() - 5.0), H_f32(10.0 * r() - 5.0)))
    elif op == 3:
        lib.b2Body_SetAwake(body, r() < 0.5)
    elif op == 4:
        lib.b2Body_Disable(body)
    elif op == 5:
        lib.b2Body_Enable(body)
    elif op == 6:
        lib.b2Body_SetType(body, int(3 * r()))
    elif op == 7:
        lib.b2DestroyBody(body)
        alive[i] = False
    elif op == 8:
        shapes = (A.ShapeId * 8)()
        n = lib.b2Body_GetShapes(body, shapes, 8)
        if n > 1:
            lib.b2DestroyShape(shapes[int(r() * n)], r() < 0.7)
    elif op == 9:
        live = [k for k in range(len(scene.joints)) if joints_alive[k]]
        if live:
            k = live[int(r() * len(live))]
            # a joint dies with either of its bodies: only destroy it while the library still knows it
            if lib.b2Joint_IsValid(scene.joints[k]):
                lib.b2DestroyJoint(scene.joints[k])
            joints_alive[k] = False
    elif op == 10:
        lib.b2Body_SetBullet(body, r() < 0.5)
    elif op == 11:
        lib.b2Body_SetFixedRotation(body, r() < 0.5)
    elif op == 12:
        shapes = (A.ShapeId * 8)()
        n = lib.b2Body_GetShapes(body, shapes, 8)
        if n > 0:
            f = A.Filter(1 << int(3 * r()), int(r() * 8) | 1, int(3 * r()) - 1)
            lib.b2Shape_SetFilter(shapes[int(r() * n)], f)
    elif op == 13:
        lib.b2Body_ApplyAngularImpulse(body, H_f32(2.0 * r() - 1.0), True)
    elif op == 14:
        lib.b2Body_EnableSleep(body, r() < 0.5)
    else:
        shapes = (A.ShapeId * 8)()
        n = lib.b2Body_GetShapes(body, shapes, 8)
        if n > 0:
            lib.b2Shape_SetFriction(shapes[0], H_f32(r()))
    return op


def H_f32(v):
    import numpy as np
    return float(np.float32(v))


def _fuzz(ref, lib, seed, frames=200, mode=None):
    lib.f2dClearLastError()
    sessions = []
    for L in (ref, lib):
        s = scenes.random_world(L, seed=seed, count=36)
        sessions.append((L, s, scenes._Lcg(31337 + seed), [True] * len(s.bodies), [True] * len(s.joints)))
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(sessions[1][1].world, mode)
    for f in range(frames):
        ops = []
        for L, s, rng, alive, joints_alive in sessions:
            n = 1 + (f % 3 == 0)
            ops.append([_mutate(L, s, rng, alive, joints_alive) for _ in range(n)] if f % 2 == 0 else [])
        assert ops[0] == ops[1]
        for L, s, *_ in sessions:
            s.step()
        d = H.diff(H.snapshot(ref, sessions[0][1].world), H.snapshot(lib, sessions[1][1].world))
        assert d == [], "seed %d frame %d after ops %s: %s" % (seed, f, ops[0], d[:6])
        assert _events(ref, sessions[0][1].world) == _events(lib, sessions[1][1].world), "seed %d frame %d: events" % (seed, f)
    assert lib.f2dGetLastError() == b""
    for L, s, *_ in sessions:
        s.destroy()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_api_calls_between_steps_emu(ref, emu, seed):
    _fuzz(ref, emu, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_api_calls_between_steps_gpu(ref, gpu, seed):
    _fuzz(ref, gpu, seed, frames=120)
