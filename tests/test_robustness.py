"""Capacities that used to be fixed: more than four chains destroyed in a row, a sensor region that
overlaps more shapes than the old 64-entry lists held, event arrays. Each scenario is run on the reference and on ours
and compared bit for bit. CPU: the host emulation of the step templates; `-m gpu`: the CUDA product."""
import ctypes as C

import pytest

import harness as H
from forge2d_b200 import _abi as A
from forge2d_b200 import scenes


def _world(lib, gravity=(0.0, -10.0)):
    wd = lib.b2DefaultWorldDef()
    wd.gravity = A.Vec2(*gravity)
    return lib.b2CreateWorld(C.byref(wd))


def _chains_session(lib, chain_count=7):
    """Creates `chain_count` chains on one static body, destroys them all (one through b2DestroyChain, the rest in
    reverse order), creates three more (id reuse is LIFO, id_pool.c:19-38) and lets a box fall on the new terrain."""
    world = _world(lib)
    bd = lib.b2DefaultBodyDef()
    ground = lib.b2CreateBody(world, C.byref(bd))
    chains = []

    def make(x0):
        pts = [(x0 + 6.0, 2.0), (x0 + 4.0, 0.0), (x0 + 2.0, 0.5), (x0, 0.0), (x0 - 2.0, 2.0)]
        arr = (A.Vec2 * len(pts))(*[A.Vec2(x, y) for x, y in pts])
        cd = lib.b2DefaultChainDef()
        cd.points = arr
        cd.count = len(pts)
        cd.isLoop = False
        return lib.b2CreateChain(ground, C.byref(cd))

    for k in range(chain_count):
        chains.append(make(-30.0 + 9.0 * k))
    ids = [c.index1 for c in chains]
    lib.b2DestroyChain(chains[2])
    for k in reversed(range(chain_count)):
        if k != 2:
            lib.b2DestroyChain(chains[k])
    chains = [make(-4.0), make(5.0), make(-13.0)]
    ids += [c.index1 for c in chains]
    bd = lib.b2DefaultBodyDef()
    bd.type = 2
    bd.position = A.Vec2(1.0, 6.0)
    box = lib.b2CreateBody(world, C.byref(bd))
    sd = lib.b2DefaultShapeDef()
    poly = lib.b2MakeBox(0.5, 0.5)
    lib.b2CreatePolygonShape(box, C.byref(sd), C.byref(poly))
    ys = []
    for _ in range(120):
        lib.b2World_Step(world, scenes.TIME_STEP, scenes.SUB_STEPS)
        ys.append(lib.b2Body_GetPosition(box).y)
    snap = H.snapshot(lib, world)
    lib.b2DestroyWorld(world)
    return ids, ys, snap


def _compare_chains(ref, lib):
    ids_a, ys_a, snap_a = _chains_session(ref)
    ids_b, ys_b, snap_b = _chains_session(lib)
    assert ids_a == ids_b, "chain ids: reference %r, ours %r" % (ids_a, ids_b)
    assert ys_a == ys_b
    assert ys_b[-1] < 5.0, "the box must have fallen (a frozen world would leave it at y = 6)"
    assert H.diff(snap_a, snap_b) == []
    assert lib.f2dGetLastError() == b""


def _sensor_session(lib, visitors=150):
    """One large static sensor box and `visitors` small circles inside it (all with sensor events, the default):
    every one of them begins to overlap in the first step; later a third of them leaves."""
    world = _world(lib, gravity=(0.0, 0.0))
    bd = lib.b2DefaultBodyDef()
    zone = lib.b2CreateBody(world, C.byref(bd))
    ssd = lib.b2DefaultShapeDef()
    ssd.isSensor = True
    ssd.enableSensorEvents = True
    box = lib.b2MakeBox(40.0, 40.0)
    lib.b2CreatePolygonShape(zone, C.byref(ssd), C.byref(box))
    sd = lib.b2DefaultShapeDef()
    sd.enableSensorEvents = True
    bodies = []
    for k in range(visitors):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(-30.0 + 4.0 * (k % 16), -30.0 + 4.0 * (k // 16))
        if k % 3 == 0:
            bd.linearVelocity = A.Vec2(60.0, 0.0)
        b = lib.b2CreateBody(world, C.byref(bd))
        c = A.Circle(A.Vec2(0.0, 0.0), 0.5)
        lib.b2CreateCircleShape(b, C.byref(sd), C.byref(c))
        bodies.append(b)
    log = []
    for _ in range(90):
        lib.b2World_Step(world, scenes.TIME_STEP, scenes.SUB_STEPS)
        ev = H.events(lib, world)
        log.append((ev["sensor_begin"], ev["sensor_end"]))
    xs = [lib.b2Body_GetPosition(b).x for b in bodies]
    lib.b2DestroyWorld(world)
    return log, xs


def _compare_sensors(ref, lib):
    log_a, xs_a = _sensor_session(ref)
    log_b, xs_b = _sensor_session(lib)
    assert len(log_a[0][0]) == 150, "the reference reports every visitor in the first step"
    assert log_a == log_b
    assert xs_a == xs_b
    assert any(len(e) > 0 for _, e in log_a), "some visitors must have left the sensor"
    assert lib.f2dGetLastError() == b""


def test_many_chains_destroyed_emu(ref, emu):
    _compare_chains(ref, emu)


def test_sensor_with_many_overlaps_emu(ref, emu):
    _compare_sensors(ref, emu)


@pytest.mark.gpu
def test_many_chains_destroyed_gpu(ref, product):
    _compare_chains(ref, product)


@pytest.mark.gpu
def test_sensor_with_many_overlaps_gpu(ref, product):
    _compare_sensors(ref, product)
