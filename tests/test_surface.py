"""The wider b2* surface (body / shape / joint accessors and mutators, destruction, all joint types) against the
compiled reference: identical scripted sessions (tests/surface_script.py) and the joint_zoo scene, every observation
and every internal record bit-identical. CPU: the host emulation of the step templates; `-m gpu`: the CUDA product."""
import pytest

import harness as H
import surface_script
from forge2d_b200 import scenes


def _compare_sessions(ref, lib):
    obs_a, snaps_a = surface_script.run(ref)
    obs_b, snaps_b = surface_script.run(lib)
    assert len(obs_a) == len(obs_b)
    for (la, va), (lb, vb) in zip(obs_a, obs_b):
        assert la == lb
        assert va == vb, "observation %s: reference %r, ours %r" % (la, va, vb)
    for (fa, sa), (fb, sb) in zip(snaps_a, snaps_b):
        d = H.diff(sa, sb)
        assert d == [], "frame %d: %s" % (fa, d[:6])


def _joint_zoo(ref, lib, frames, every, mode=None):
    lib.f2dClearLastError()
    a = scenes.joint_zoo(ref, sets=3)
    b = scenes.joint_zoo(lib, sets=3)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    for f in range(frames):
        a.step()
        b.step()
        if f % every == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
            assert d == [], "frame %d: %s" % (f, d[:6])
            for ja, jb in zip(a.joints, b.joints):
                fa, fb = ref.b2Joint_GetConstraintForce(ja), lib.b2Joint_GetConstraintForce(jb)
                assert (fa.x, fa.y) == (fb.x, fb.y)
                assert ref.b2Joint_GetConstraintTorque(ja) == lib.b2Joint_GetConstraintTorque(jb)
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_scripted_session_matches_reference_emu(ref, emu):
    _compare_sessions(ref, emu)


def test_all_joint_types_bit_identical_emu(ref, emu):
    _joint_zoo(ref, emu, 300, 5)


# ---- host callbacks inside the step -------------------------------------------------------------------------
def _callback_session(ref, lib, frames=150, mode=None):
    """b2CustomFilterFcn and b2PreSolveFcn (world.c:1710-1740) registered on a falling-boxes world: the custom filter
    rejects pairs by shape-id parity, the pre-solve callback makes a platform one-way (contacts whose normal points
    down are disabled) and rejects by manifold content. Same verdicts given => every record bit-identical, and the
    callbacks see the same (shape ids, manifold) sequence as in the single-worker reference."""
    import ctypes as C
    from forge2d_b200 import _abi as A
    FilterFcn = C.CFUNCTYPE(C.c_bool, A.ShapeId, A.ShapeId, C.c_void_p)
    PreSolveFcn = C.CFUNCTYPE(C.c_bool, A.ShapeId, A.ShapeId, C.POINTER(A.Manifold), C.c_void_p)
    lib.f2dClearLastError()
    logs, keep, worlds = [], [], []
    for L in (ref, lib):
        s = scenes.bench2d(L, rows=9, ground_half_width=12.0)
        sd = L.b2DefaultShapeDef()
        sd.enablePreSolveEvents = True
        bd = L.b2DefaultBodyDef()
        bd.position = A.Vec2(0.0, -24.0)
        platform = L.b2CreateBody(s.world, C.byref(bd))
        box = L.b2MakeBox(6.0, 0.25)
        L.b2CreatePolygonShape(platform, C.byref(sd), C.byref(box))
        bd.type = 2
        for k in range(10):     # boxes above and below the platform; those below are thrown upwards through it
            bd.position = A.Vec2(-4.5 + k, -21.0 if k % 2 else -27.5)
            bd.linearVelocity = A.Vec2(0.0, 0.0 if k % 2 else 14.0)
            b = L.b2CreateBody(s.world, C.byref(bd))
            small = L.b2MakeBox(0.3, 0.3)
            L.b2CreatePolygonShape(b, C.byref(sd), C.byref(small))
        log = []

        def custom_filter(a, b, ctx, log=log):
            log.append(("filter", a.index1, a.generation, b.index1, b.generation))
            return (a.index1 + b.index1) % 5 != 0

        def pre_solve(a, b, manifold, ctx, log=log):
            m = manifold.contents
            log.append(("presolve", a.index1, b.index1, m.pointCount, m.normal.x, m.normal.y,
                        tuple((m.points[i].id, m.points[i].separation, m.points[i].anchorA.x, m.points[i].normalImpulse)
                              for i in range(m.pointCount))))
            return m.normal.y > -0.5 and (a.index1 * 7 + b.index1) % 11 != 0

        f1, f2 = FilterFcn(custom_filter), PreSolveFcn(pre_solve)
        keep += [f1, f2]
        L.b2World_SetCustomFilterCallback(s.world, C.cast(f1, C.c_void_p), None)
        L.b2World_SetPreSolveCallback(s.world, C.cast(f2, C.c_void_p), None)
        logs.append(log)
        worlds.append(s)
    a, b = worlds
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    for f in range(frames):
        if f == 100:   # callbacks can be removed again: back to the single-launch step
            for s, L in ((a, ref), (b, lib)):
                L.b2World_SetPreSolveCallback(s.world, None, None)
        a.step()
        b.step()
        d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
        assert d == [], "frame %d: %s" % (f, d[:6])
        assert logs[0] == logs[1], "frame %d: callback sequences differ" % f
    kinds = [e[0] for e in logs[0]]
    assert kinds.count("filter") > 100 and kinds.count("presolve") > 100
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_in_step_host_callbacks_match_reference_emu(ref, emu):
    _callback_session(ref, emu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_in_step_host_callbacks_match_reference_gpu(ref, gpu, mode):
    _callback_session(ref, gpu, mode=mode)


@pytest.mark.gpu
def test_scripted_session_matches_reference_gpu(ref, gpu):
    _compare_sessions(ref, gpu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_all_joint_types_bit_identical_gpu(ref, gpu, mode):
    _joint_zoo(ref, gpu, 300, 5, mode)


# ---- queries -------------------------------------------------------------------------------------------------
def _query_session(lib, frames=90):
    """Ray casts (all hits in callback order, clipped, closest), AABB overlaps and an explosion on a mixed scene."""
    import ctypes as C
    from forge2d_b200 import _abi as A
    s = scenes.falling_shapes(lib, count=24) if False else scenes.bench2d(lib, rows=10, ground_half_width=14.0)
    world = s.world
    sd = lib.b2DefaultShapeDef()
    bd = lib.b2DefaultBodyDef()
    bd.type = 2
    for k, (x, y) in enumerate(((-9.0, -20.0), (-8.0, -12.0), (9.0, -18.0), (5.0, -15.0))):
        bd.position = A.Vec2(x, y)
        b = lib.b2CreateBody(world, C.byref(bd))
        if k == 0:
            c = A.Circle(A.Vec2(0.0, 0.0), 0.6)
            lib.b2CreateCircleShape(b, C.byref(sd), C.byref(c))
        elif k == 1:
            c = A.Capsule(A.Vec2(-0.5, 0.0), A.Vec2(0.5, 0.2), 0.3)
            lib.b2CreateCapsuleShape(b, C.byref(sd), C.byref(c))
        elif k == 3:
            # rounded polygon: ray casts go through b2ShapeCast (geometry.c:877-887)
            rounded = lib.b2MakeOffsetRoundedBox(0.5, 0.3, A.Vec2(0.1, 0.0), A.Rot(0.8, 0.6), 0.25)
            lib.b2CreatePolygonShape(b, C.byref(sd), C.byref(rounded))
        else:
            sd2 = lib.b2DefaultShapeDef()
            sd2.filter.categoryBits = 2
            box = lib.b2MakeBox(0.7, 0.3)
            lib.b2CreatePolygonShape(b, C.byref(sd2), C.byref(box))
    out = []

    def bits(x):
        import numpy as np
        return int(np.float32(x).view(np.uint32))

    rays = [((-20.0, -29.5), (40.0, 0.0)), ((-6.0, 5.0), (3.0, -40.0)), ((0.0, -10.0), (0.0, -25.0)),
            ((-12.0, -25.0), (30.0, 9.0)), ((2.0, -28.0), (0.0, 0.0)), ((-9.0, -20.0), (1.0, 1.0)),
            ((5.15, -10.0), (0.0, -22.0)), ((-2.0, -14.0), (14.0, -6.0)), ((5.0, -15.3), (1.0, 0.0)), ((8.0, -31.0), (-4.5, 17.0))]
    for f in range(frames):
        s.step()
        if f % 15 != 14:
            continue
        flt = lib.b2DefaultQueryFilter()
        for ri, (o, t) in enumerate(rays):
            for mode in ("all", "clip", "first"):
                hits = []

                def cb(shape, point, normal, fraction, ctx, hits=hits, mode=mode):
                    hits.append((shape.index1, bits(point.x), bits(point.y), bits(normal.x), bits(normal.y), bits(fraction)))
                    return {"all": 1.0, "clip": fraction, "first": 0.0}[mode]

                st = lib.b2World_CastRay(world, A.Vec2(*o), A.Vec2(*t), flt, A.CastResultFcn(cb), None)
                out.append(("ray%d.%s.f%d" % (ri, mode, f), hits, st.nodeVisits, st.leafVisits))
            r = lib.b2World_CastRayClosest(world, A.Vec2(*o), A.Vec2(*t), flt)
            out.append(("closest%d.f%d" % (ri, f), r.hit, r.shapeId.index1 if r.hit else 0, bits(r.fraction), bits(r.point.x),
                        bits(r.normal.y), r.nodeVisits, r.leafVisits))
        flt2 = lib.b2DefaultQueryFilter()
        flt2.maskBits = 2
        for box, fl in ((A.AABB(A.Vec2(-3.0, -30.5), A.Vec2(3.0, -27.0)), flt), (A.AABB(A.Vec2(-20.0, -31.0), A.Vec2(20.0, 0.0)), flt2)):
            found = []

            def ocb(shape, ctx, found=found):
                found.append(shape.index1)
                return len(found) < 25

            st = lib.b2World_OverlapAABB(world, box, fl, A.OverlapResultFcn(ocb), None)
            out.append(("overlap.f%d" % f, found, st.nodeVisits, st.leafVisits))
        if f == 44:
            ex = lib.b2DefaultExplosionDef()
            ex.position = A.Vec2(0.0, -29.0)
            ex.radius, ex.falloff, ex.impulsePerLength = 3.0, 2.0, 8.0
            lib.b2World_Explode(world, C.byref(ex))
    snap = H.snapshot(lib, world)
    s.destroy()
    return out, snap


def test_queries_match_reference_emu(ref, emu):
    a, sa = _query_session(ref)
    b, sb = _query_session(emu)
    assert a == b
    assert H.diff(sa, sb) == []


@pytest.mark.gpu
def test_queries_match_reference_gpu(ref, gpu):
    a, sa = _query_session(ref)
    b, sb = _query_session(gpu)
    assert a == b
    assert H.diff(sa, sb) == []


# ---- sensors -------------------------------------------------------------------------------------------------
def _sensor_session(ref, lib, frames=260, mode=None):
    a = scenes.sensor_field(ref)
    b = scenes.sensor_field(lib)
    lib.f2dClearLastError()
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    seen = 0
    for f in range(frames):
        if f == 90:      # a visitor and a sensor shape disappear mid-run; a visitor opts out
            for s, L in ((a, ref), (b, lib)):
                L.b2DestroyBody(s.bodies[9])
                L.b2DestroyShape(s.sensors[0], False)
                import ctypes as C
                from forge2d_b200 import _abi as A
                arr = (A.ShapeId * 2)()
                L.b2Body_GetShapes(s.bodies[12], arr, 2)
                L.b2Shape_EnableSensorEvents(arr[0], False)
        a.step()
        b.step()
        ea, eb = H.events(ref, a.world), H.events(lib, b.world)
        assert ea["sensor_begin"] == eb["sensor_begin"], "sensor begin events, frame %d" % f
        assert ea["sensor_end"] == eb["sensor_end"], "sensor end events, frame %d" % f
        seen += len(ea["sensor_begin"]) + len(ea["sensor_end"])
        if f % 20 == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
            assert d == [], "frame %d: %s" % (f, d[:6])
    assert seen > 40, "the scene is supposed to produce sensor traffic (saw %d events)" % seen
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_sensor_events_match_reference_emu(ref, emu):
    _sensor_session(ref, emu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_sensor_events_match_reference_gpu(ref, gpu, mode):
    _sensor_session(ref, gpu, mode=mode)


# ---- chains --------------------------------------------------------------------------------------------------
def _chain_session(ref, lib, frames=300, mode=None):
    import ctypes as C
    from forge2d_b200 import _abi as A
    lib.f2dClearLastError()
    a = scenes.chain_terrain(ref)
    b = scenes.chain_terrain(lib)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    for s, L in ((a, ref), (b, lib)):
        assert L.b2Chain_IsValid(s.chains[0]) and L.b2Chain_GetSegmentCount(s.chains[0]) == 11
        assert L.b2Chain_GetSegmentCount(s.chains[1]) == 5
    segs_a, segs_b = (A.ShapeId * 16)(), (A.ShapeId * 16)()
    assert ref.b2Chain_GetSegments(a.chains[1], segs_a, 16) == lib.b2Chain_GetSegments(b.chains[1], segs_b, 16)
    assert [segs_a[i].index1 for i in range(5)] == [segs_b[i].index1 for i in range(5)]
    assert ref.b2Chain_GetFriction(a.chains[1]) == lib.b2Chain_GetFriction(b.chains[1])
    for f in range(frames):
        if f == 120:
            for s, L in ((a, ref), (b, lib)):
                L.b2Chain_SetFriction(s.chains[0], 0.05)
                L.b2Chain_SetRestitution(s.chains[0], 0.4)
        if f == 200:
            for s, L in ((a, ref), (b, lib)):
                L.b2DestroyChain(s.chains[1])     # the basin vanishes, its contents fall onto the terrain
                assert not L.b2Chain_IsValid(s.chains[1])
        a.step()
        b.step()
        if f % 6 == 0 or f == frames - 1:
            d = H.diff(H.snapshot(ref, a.world), H.snapshot(lib, b.world))
            assert d == [], "frame %d: %s" % (f, d[:6])
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_chain_shapes_match_reference_emu(ref, emu):
    _chain_session(ref, emu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_chain_shapes_match_reference_gpu(ref, gpu, mode):
    _chain_session(ref, gpu, mode=mode)


# ---- debug draw ----------------------------------------------------------------------------------------------
DRAW_OPTION_SETS = [
    dict(drawShapes=True),
    dict(drawShapes=True, drawJoints=True, drawJointExtras=True, drawBounds=True, drawMass=True, drawBodyNames=True,
         drawContacts=True, drawContactNormals=True, drawContactFeatures=True, drawFrictionImpulses=True, drawIslands=True),
    dict(drawJoints=True, drawGraphColors=True, drawContacts=True, drawContactImpulses=True),
]


def _draw_session(ref, lib, make, frames, mode=None, bounds=(-6.0, -2.0, 9.0, 14.0)):
    """b2World_Draw emits the same primitives, in the same order, with the same bits as the reference (world.c:1161-1489),
    for the whole world and for a drawing window, on freshly built and on stepped (partly sleeping) worlds."""
    lib.f2dClearLastError()
    a, b = make(ref), make(lib)
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b.world, mode)
    for s, L in ((a, ref), (b, lib)):
        for k, body in enumerate(s.bodies[:12:3]):
            L.b2Body_SetName(body, b"body-%d" % k)
    emitted = 0
    for f in range(frames + 1):
        if f in (0, 3, frames // 2, frames):
            for options in DRAW_OPTION_SETS:
                for window in (None, bounds):
                    ca = H.record_draw(ref, a.world, bounds=window, **options)
                    cb = H.record_draw(lib, b.world, bounds=window, **options)
                    assert len(ca) == len(cb), "frame %d %s window %s: %d vs %d primitives" % (f, sorted(options), window, len(ca), len(cb))
                    for i, (x, y) in enumerate(zip(ca, cb)):
                        assert x == y, "frame %d %s window %s, primitive %d: reference %r, ours %r" % (f, sorted(options), window, i, x, y)
                    emitted += len(ca)
        a.step()
        b.step()
    assert emitted > 1000
    assert lib.f2dGetLastError() == b""
    a.destroy()
    b.destroy()


def test_debug_draw_matches_reference_emu(ref, emu):
    _draw_session(ref, emu, lambda L: scenes.joint_zoo(L, sets=2), 90)
    _draw_session(ref, emu, scenes.chain_terrain, 60)
    _draw_session(ref, emu, scenes.sensor_field, 40)
    _draw_session(ref, emu, lambda L: scenes.many_pyramids(L, grid=3, base=4), 120)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_debug_draw_matches_reference_gpu(ref, gpu, mode):
    _draw_session(ref, gpu, lambda L: scenes.joint_zoo(L, sets=2), 90, mode=mode)
    _draw_session(ref, gpu, scenes.chain_terrain, 60, mode=mode)
    _draw_session(ref, gpu, lambda L: scenes.many_pyramids(L, grid=3, base=4), 120, mode=mode)


# ---- joints between sleeping bodies ----------------------------------------------------------------------------
def _sleeping_joint_session(ref, lib, mode=None):
    """Joints created between bodies that are asleep: the joint goes into the sleeping solver set (joint.c:250-287),
    two different sleeping sets are merged (solver_set.c:427-519); joints follow a disabled / re-enabled / retyped
    body into sleeping sets (body.c). Every record equal after each edit and while the chains wake up and swing."""
    import ctypes as C
    from forge2d_b200 import _abi as A
    lib.f2dClearLastError()
    def make(L):
        s = scenes.bench2d(L, rows=3, ground_half_width=20.0)
        world = s.world
        sd = L.b2DefaultShapeDef()
        box = L.b2MakeBox(0.4, 0.2)

        def body(x, y, kind=2, awake=False):
            bd = L.b2DefaultBodyDef()
            bd.type = kind
            bd.position = A.Vec2(x, y)
            bd.isAwake = awake
            b = L.b2CreateBody(world, C.byref(bd))
            L.b2CreatePolygonShape(b, C.byref(sd), C.byref(box))
            return b

        sleepers = [body(-8.0 + 1.5 * k, -20.0 + 0.1 * k) for k in range(7)]
        anchor = body(-9.5, -20.0, kind=0)
        joints = []

        def revolute(a, b, ax, ay):
            d = L.b2DefaultRevoluteJointDef()
            d.bodyIdA, d.bodyIdB = a, b
            d.localAnchorA, d.localAnchorB = A.Vec2(ax, ay), A.Vec2(-ax, ay)
            joints.append(L.b2CreateRevoluteJoint(world, C.byref(d)))

        def distance(a, b, length):
            d = L.b2DefaultDistanceJointDef()
            d.bodyIdA, d.bodyIdB = a, b
            d.length = length
            joints.append(L.b2CreateDistanceJoint(world, C.byref(d)))

        return dict(L=L, s=s, sleepers=sleepers, anchor=anchor, joints=joints, revolute=revolute, distance=distance)

    sessions = [make(ref), make(lib)]

    def compare(label):
        a, b = sessions
        d = H.diff(H.snapshot(ref, a["s"].world), H.snapshot(lib, b["s"].world))
        assert d == [], "%s: %s" % (label, d[:6])

    def both(fn):
        for x in sessions:
            fn(x)

    a, b = sessions
    if mode is not None:
        lib.f2dWorld_SetLaunchMode(b["s"].world, mode)
    compare("created asleep")
    both(lambda x: x["revolute"](x["sleepers"][0], x["sleepers"][1], 0.75, 0.0))    # two one-body sleeping sets merge
    compare("joint between two sleeping sets")
    both(lambda x: x["distance"](x["anchor"], x["sleepers"][0], 1.5))               # static + sleeping
    compare("joint static-sleeping")
    both(lambda x: x["revolute"](x["sleepers"][2], x["sleepers"][3], 0.75, 0.0))
    both(lambda x: x["revolute"](x["sleepers"][3], x["sleepers"][4], 0.75, 0.0))    # grows the merged set again
    both(lambda x: x["revolute"](x["sleepers"][1], x["sleepers"][2], 0.75, 0.0))    # merges the two chains (2 + 3 bodies)
    compare("chains merged")
    for f in range(5):
        both(lambda x: x["s"].step())
    compare("stepped while asleep")
    both(lambda x: x["L"].b2Body_Disable(x["anchor"]))
    compare("anchor disabled")
    both(lambda x: x["L"].b2Body_Enable(x["anchor"]))                               # its joint returns to the sleeping set
    compare("anchor enabled")
    both(lambda x: x["L"].b2Body_SetType(x["anchor"], 1))
    both(lambda x: x["L"].b2Body_SetType(x["anchor"], 0))
    compare("anchor retyped")
    both(lambda x: x["L"].b2DestroyJoint(x["joints"][2]))
    compare("joint destroyed inside the sleeping set")
    both(lambda x: x["revolute"](x["sleepers"][5], x["sleepers"][6], 0.75, 0.0))
    both(lambda x: x["L"].b2Body_SetAwake(x["sleepers"][5], True))
    for f in range(40):
        both(lambda x: x["s"].step())
        if f % 8 == 0:
            compare("frame %d after waking one pair" % f)
    both(lambda x: x["L"].b2Body_SetAwake(x["sleepers"][1], True))
    for f in range(120):
        both(lambda x: x["s"].step())
        if f % 8 == 0 or f == 119:
            compare("frame %d after waking the chain" % f)
    assert lib.f2dGetLastError() == b""
    both(lambda x: x["s"].destroy())


def test_joints_between_sleeping_bodies_match_reference_emu(ref, emu):
    _sleeping_joint_session(ref, emu)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1], ids=["cta", "grid"])
def test_joints_between_sleeping_bodies_match_reference_gpu(ref, gpu, mode):
    _sleeping_joint_session(ref, gpu, mode=mode)
