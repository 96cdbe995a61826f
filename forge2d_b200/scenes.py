"""Synthetic scene generators for the five BASELINE.json configurations.

Each builder issues the SAME b2* call sequence against whatever library it is handed (product, host emulation or
the compiled reference), which is what makes bit-level parity comparisons meaningful.
"""
import ctypes as C
from . import _abi as A

TIME_STEP = 1.0 / 60.0
SUB_STEPS = 4


class Scene:
    def __init__(self, lib, world, bodies, name):
        self.lib, self.world, self.bodies, self.name = lib, world, bodies, name
        self.frame = 0

    def step(self, dt=TIME_STEP, sub=SUB_STEPS):
        self.lib.b2World_Step(self.world, dt, sub)
        self.frame += 1

    def destroy(self):
        self.lib.b2DestroyWorld(self.world)


def _world(lib, enable_sleep=True, enable_continuous=True, gravity=(0.0, -10.0), create=None):
    wd = lib.b2DefaultWorldDef()
    wd.gravity = A.Vec2(*gravity)
    wd.enableSleep = enable_sleep
    wd.enableContinuous = enable_continuous
    if create is not None:
        return create(wd)
    return lib.b2CreateWorld(C.byref(wd))


def _static_segment(lib, world, p1, p2, origin=(0.0, 0.0)):
    bd = lib.b2DefaultBodyDef()
    bd.position = A.Vec2(*origin)
    ground = lib.b2CreateBody(world, C.byref(bd))
    sd = lib.b2DefaultShapeDef()
    seg = A.Segment(A.Vec2(*p1), A.Vec2(*p2))
    lib.b2CreateSegmentShape(ground, C.byref(sd), C.byref(seg))
    return ground


def _box(lib, half):
    # Dart's Polygon.square/box goes through b2MakeOffsetRoundedBox (raw_box2d_ffi.dart:682-690)
    return lib.b2MakeOffsetRoundedBox(half, half, A.Vec2(0.0, 0.0), A.Rot(1.0, 0.0), 0.0)


def bench2d(lib, rows=40, ground_half_width=40.0, create=None, x_offset=0.0, **world_kw):
    """C1: packages/benchmark/bin/bench2d.dart:29-51 — `rows`-high pyramid of 0.5 half-extent boxes, density 5.
    `x_offset` builds the same scene translated along x (decorrelated replicas: same physics, different rounding)."""
    world = _world(lib, create=create, **world_kw)
    bodies = [_static_segment(lib, world, (_f32(x_offset - ground_half_width), -30.0), (_f32(x_offset + ground_half_width), -30.0))]
    box = _box(lib, 0.5)
    sd = lib.b2DefaultShapeDef()
    sd.density = 5.0
    x = [_f32(-7.0 + x_offset), 0.75]
    dxx, dxy = 0.5625, 1.0
    for i in range(rows):
        y = list(x)
        for _j in range(i, rows):
            bd = lib.b2DefaultBodyDef()
            bd.type = 2
            bd.position = A.Vec2(_f32(y[0]), _f32(y[1]))
            body = lib.b2CreateBody(world, C.byref(bd))
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(box))
            bodies.append(body)
            y[0] = _f32(y[0] + 1.125)
        x[0] = _f32(x[0] + dxx)
        x[1] = _f32(x[1] + dxy)
    return Scene(lib, world, bodies, "bench2d_%d" % rows)


def _f32(v):
    return C.c_float(v).value


def large_pyramid(lib, rows=100, create=None, **world_kw):
    """C2: the bench2d generator with 100 rows (5050 bodies) on a ground widened to +-200."""
    s = bench2d(lib, rows=rows, ground_half_width=200.0, create=create, **world_kw)
    s.name = "large_pyramid_%d" % rows
    return s


def many_pyramids(lib, grid=20, base=10, create=None, **world_kw):
    """C3: grid x grid small pyramids (base `base`, boxes of half-extent 0.5 stacked touching), one ground segment
    per row; upstream Box2D's many_pyramids layout. grid=20, base=10 -> 22 000 bodies, 400 islands."""
    world = _world(lib, create=create, **world_kw)
    bodies = []
    extent = 0.5
    box = _box(lib, extent)
    sd = lib.b2DefaultShapeDef()
    bd0 = lib.b2DefaultBodyDef()
    ground = lib.b2CreateBody(world, C.byref(bd0))
    bodies.append(ground)
    base_width = 2.0 * extent * base
    base_y = 0.0
    delta_x = 2.0 * extent * (base + 1.0)
    delta_y = 2.0 * extent * (base + 1.0) + 5.0 * extent
    total_width = delta_x * grid
    for i in range(grid):
        seg = A.Segment(A.Vec2(_f32(-0.5 * total_width - base_width), _f32(base_y + i * delta_y)),
                        A.Vec2(_f32(0.5 * total_width + base_width), _f32(base_y + i * delta_y)))
        lib.b2CreateSegmentShape(ground, C.byref(sd), C.byref(seg))
    for i in range(grid):
        row_y = base_y + i * delta_y
        for j in range(grid):
            center_x = -0.5 * total_width + j * delta_x + extent * base
            for r in range(base):
                y = (2.0 * r + 1.0) * extent + row_y
                for c in range(r, base):
                    x = (r + 1.0) * extent + 2.0 * (c - r) * extent + center_x - extent * base - 0.5
                    bd = lib.b2DefaultBodyDef()
                    bd.type = 2
                    bd.position = A.Vec2(_f32(x), _f32(y))
                    body = lib.b2CreateBody(world, C.byref(bd))
                    lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(box))
                    bodies.append(body)
    return Scene(lib, world, bodies, "many_pyramids_%dx%d_b%d" % (grid, grid, base))


class _Lcg:
    """Fixed LCG so the rain is identical for every library (seed 12345)."""

    def __init__(self, seed=12345):
        self.s = seed

    def next(self):
        self.s = (1103515245 * self.s + 12345) & 0x7FFFFFFF
        return self.s / float(0x7FFFFFFF)


class JointGrid(Scene):
    def __init__(self, lib, world, bodies, name, rain_every):
        super().__init__(lib, world, bodies, name)
        self.rng = _Lcg()
        self.rain_every = rain_every
        self.n = 0

    def step(self, dt=TIME_STEP, sub=SUB_STEPS):
        if self.rain_every > 0 and self.frame % self.rain_every == 0:
            self._rain()
        super().step(dt, sub)

    def _rain(self):
        lib, world = self.lib, self.world
        sd = lib.b2DefaultShapeDef()
        sd.filter.categoryBits = 1
        sd.filter.maskBits = 0xFFFFFFFFFFFFFFFF
        x0 = self.n * 0.5 * self.rng.next()
        for kind in range(3):
            bd = lib.b2DefaultBodyDef()
            bd.type = 2
            bd.position = A.Vec2(_f32(x0 + 10.0 + 15.0 * kind + 5.0 * self.rng.next()), _f32(5.0 + 2.0 * kind))
            body = lib.b2CreateBody(world, C.byref(bd))
            if kind == 0:
                c = A.Circle(A.Vec2(0.0, 0.0), 0.25)
                lib.b2CreateCircleShape(body, C.byref(sd), C.byref(c))
            elif kind == 1:
                cap = A.Capsule(A.Vec2(-0.25, 0.0), A.Vec2(0.25, 0.0), 0.2)
                lib.b2CreateCapsuleShape(body, C.byref(sd), C.byref(cap))
            else:
                b = _box(lib, 0.3)
                lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(b))
            self.bodies.append(body)
        self.n += 1


def joint_grid(lib, n=100, rain_every=4, create=None, **world_kw):
    """C4: n x n circles (r=0.4) on an integer grid linked to the up/left neighbour by revolute joints, 7 static
    anchors in the top row, sleeping disabled; plus a rain of circle/capsule/box every `rain_every` frames."""
    world_kw.setdefault("enable_sleep", False)
    world = _world(lib, create=create, **world_kw)
    bodies = []
    sd = lib.b2DefaultShapeDef()
    sd.filter.categoryBits = 2
    sd.filter.maskBits = 0xFFFFFFFFFFFFFFFF & ~2
    circle = A.Circle(A.Vec2(0.0, 0.0), 0.4)
    jd = lib.b2DefaultRevoluteJointDef()
    grid = [[None] * n for _ in range(n)]
    lo, hi = n // 2 - 3, n // 2 + 3
    for k in range(n):
        for i in range(n):
            bd = lib.b2DefaultBodyDef()
            if lo <= k <= hi and i == 0:
                bd.type = 0
            else:
                bd.type = 2
            bd.position = A.Vec2(_f32(float(k)), _f32(float(-i)))
            body = lib.b2CreateBody(world, C.byref(bd))
            lib.b2CreateCircleShape(body, C.byref(sd), C.byref(circle))
            grid[k][i] = body
            bodies.append(body)
            if i > 0:
                jd.bodyIdA = grid[k][i - 1]
                jd.bodyIdB = body
                jd.localAnchorA = A.Vec2(0.0, -0.5)
                jd.localAnchorB = A.Vec2(0.0, 0.5)
                lib.b2CreateRevoluteJoint(world, C.byref(jd))
            if k > 0:
                jd.bodyIdA = grid[k - 1][i]
                jd.bodyIdB = body
                jd.localAnchorA = A.Vec2(0.5, 0.0)
                jd.localAnchorB = A.Vec2(-0.5, 0.0)
                lib.b2CreateRevoluteJoint(world, C.byref(jd))
    return JointGrid(lib, world, bodies, "joint_grid_%d" % n, rain_every)


def falling_shapes(lib, count=24, create=None, **world_kw):
    """Small mixed-shape scene (circles, capsules, boxes, rounded boxes on a segment + a static box) used by the
    parity tests to cover every manifold function on the device path."""
    world = _world(lib, create=create, **world_kw)
    bodies = [_static_segment(lib, world, (-30.0, 0.0), (30.0, 0.0))]
    bd = lib.b2DefaultBodyDef()
    bd.position = A.Vec2(4.0, 1.0)
    wall = lib.b2CreateBody(world, C.byref(bd))
    sd = lib.b2DefaultShapeDef()
    wbox = lib.b2MakeBox(0.5, 1.0)
    lib.b2CreatePolygonShape(wall, C.byref(sd), C.byref(wbox))
    bodies.append(wall)
    rng = _Lcg(777)
    sd.material.restitution = 0.3
    for i in range(count):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(_f32(-6.0 + 12.0 * rng.next()), _f32(2.0 + 1.2 * i))
        bd.angularVelocity = _f32(2.0 * rng.next() - 1.0)
        body = lib.b2CreateBody(world, C.byref(bd))
        kind = i % 4
        if kind == 0:
            c = A.Circle(A.Vec2(0.0, 0.0), _f32(0.3 + 0.2 * rng.next()))
            lib.b2CreateCircleShape(body, C.byref(sd), C.byref(c))
        elif kind == 1:
            cap = A.Capsule(A.Vec2(-0.4, 0.0), A.Vec2(0.4, 0.1), _f32(0.2 + 0.1 * rng.next()))
            lib.b2CreateCapsuleShape(body, C.byref(sd), C.byref(cap))
        elif kind == 2:
            b = _box(lib, _f32(0.3 + 0.2 * rng.next()))
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(b))
        else:
            b = lib.b2MakeOffsetRoundedBox(0.3, 0.2, A.Vec2(0.0, 0.0), A.Rot(1.0, 0.0), 0.1)
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(b))
        bodies.append(body)
    return Scene(lib, world, bodies, "falling_shapes_%d" % count)


def polygon_soup(lib, count=40, create=None, **world_kw):
    """Convex polygons with 3..8 vertices (b2ComputeHull + b2MakePolygon, some rounded), capsules and circles dropped
    into a box of segments: covers the 8-vertex SAT / clipping path and every polygon-vs-other manifold function."""
    import math
    world = _world(lib, create=create, **world_kw)
    bd = lib.b2DefaultBodyDef()
    ground = lib.b2CreateBody(world, C.byref(bd))
    sd = lib.b2DefaultShapeDef()
    for p1, p2 in (((-12.0, 0.0), (12.0, 0.0)), ((-12.0, 0.0), (-14.0, 20.0)), ((12.0, 0.0), (14.0, 20.0))):
        seg = A.Segment(A.Vec2(*p1), A.Vec2(*p2))
        lib.b2CreateSegmentShape(ground, C.byref(sd), C.byref(seg))
    bodies = [ground]
    rng = _Lcg(4242)
    sd.material.friction = 0.4
    for i in range(count):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(_f32(-8.0 + 16.0 * rng.next()), _f32(1.5 + 0.9 * i))
        ang = 6.28 * rng.next()
        bd.rotation = A.Rot(_f32(math.cos(ang)), _f32(math.sin(ang)))
        bd.angularVelocity = _f32(4.0 * rng.next() - 2.0)
        body = lib.b2CreateBody(world, C.byref(bd))
        kind = i % 5
        if kind == 3:
            cap = A.Capsule(A.Vec2(-0.5, 0.0), A.Vec2(0.5, 0.0), _f32(0.15 + 0.2 * rng.next()))
            lib.b2CreateCapsuleShape(body, C.byref(sd), C.byref(cap))
        elif kind == 4:
            c = A.Circle(A.Vec2(0.0, 0.0), _f32(0.25 + 0.3 * rng.next()))
            lib.b2CreateCircleShape(body, C.byref(sd), C.byref(c))
        else:
            n = 3 + (i * 7) % 6  # 3..8 vertices
            r = 0.4 + 0.4 * rng.next()
            pts = (A.Vec2 * n)()
            for k in range(n):
                a = 2.0 * math.pi * (k + 0.3 * rng.next()) / n
                pts[k] = A.Vec2(_f32(r * math.cos(a)), _f32(0.8 * r * math.sin(a)))
            hull = lib.b2ComputeHull(pts, n)
            poly = lib.b2MakePolygon(C.byref(hull), _f32(0.05 if kind == 2 else 0.0))
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(poly))
        bodies.append(body)
    return Scene(lib, world, bodies, "polygon_soup_%d" % count)


def jointed_piles(lib, chains=6, links=4, create=None, **world_kw):
    """Chains of boxes linked by revolute joints dropped next to loose boxes, sleeping enabled: islands that contain
    joints merge, lose contacts, get split (joint edges in the island DFS) and fall asleep."""
    world = _world(lib, create=create, **world_kw)
    bodies = [_static_segment(lib, world, (-40.0, 0.0), (40.0, 0.0))]
    sd = lib.b2DefaultShapeDef()
    box = _box(lib, 0.25)
    jd = lib.b2DefaultRevoluteJointDef()
    for c in range(chains):
        prev = None
        x0 = -12.0 + 4.5 * c
        for k in range(links):
            bd = lib.b2DefaultBodyDef()
            bd.type = 2
            bd.position = A.Vec2(_f32(x0 + 0.55 * k), _f32(1.0 + 0.3 * c))
            body = lib.b2CreateBody(world, C.byref(bd))
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(box))
            bodies.append(body)
            if prev is not None:
                jd.bodyIdA, jd.bodyIdB = prev, body
                jd.localAnchorA = A.Vec2(0.275, 0.0)
                jd.localAnchorB = A.Vec2(-0.275, 0.0)
                lib.b2CreateRevoluteJoint(world, C.byref(jd))
            prev = body
        for k in range(3):
            bd = lib.b2DefaultBodyDef()
            bd.type = 2
            bd.position = A.Vec2(_f32(x0 + 0.4 * k), _f32(2.5 + 0.7 * k))
            body = lib.b2CreateBody(world, C.byref(bd))
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(box))
            bodies.append(body)
    return Scene(lib, world, bodies, "jointed_piles_%d" % chains)


def joint_zoo(lib, sets=3, create=None, **world_kw):
    """Every joint type with its options exercised (springs, limits, motors, soft welds): `sets` copies side by side,
    each a static anchor bar plus dynamic boxes / circles hanging, sliding, rolling or being dragged from it."""
    world = _world(lib, create=create, **world_kw)
    bodies = [_static_segment(lib, world, (-60.0, -6.0), (60.0, -6.0))]
    joints = []
    sd = lib.b2DefaultShapeDef()
    box = _box(lib, 0.3)
    wheel = A.Circle(A.Vec2(0.0, 0.0), 0.35)

    def dyn(x, y, shape="box", angle=0.0):
        import math
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(_f32(x), _f32(y))
        bd.rotation = A.Rot(_f32(math.cos(angle)), _f32(math.sin(angle)))
        b = lib.b2CreateBody(world, C.byref(bd))
        if shape == "box":
            lib.b2CreatePolygonShape(b, C.byref(sd), C.byref(box))
        else:
            lib.b2CreateCircleShape(b, C.byref(sd), C.byref(wheel))
        bodies.append(b)
        return b

    for k in range(sets):
        x0 = -30.0 + 20.0 * k
        bd = lib.b2DefaultBodyDef()
        bd.position = A.Vec2(_f32(x0), 4.0)
        anchor = lib.b2CreateBody(world, C.byref(bd))
        bodies.append(anchor)
        # distance: rigid rod, spring with limits, motorised spring
        for i, (spring, limit, motor) in enumerate(((False, False, False), (True, True, False), (True, False, True))):
            b = dyn(x0 - 6.0 + 1.5 * i, 2.0 - 0.2 * k)
            d = lib.b2DefaultDistanceJointDef()
            d.bodyIdA, d.bodyIdB = anchor, b
            d.localAnchorA = A.Vec2(_f32(-6.0 + 1.5 * i), 0.0)
            d.localAnchorB = A.Vec2(0.0, 0.25)
            d.length = _f32(1.75 + 0.1 * k)
            d.enableSpring, d.hertz, d.dampingRatio = spring, 2.0, 0.3
            d.enableLimit, d.minLength, d.maxLength = limit, 1.5, 2.25
            d.enableMotor, d.maxMotorForce, d.motorSpeed = motor, 40.0, _f32(0.5 - 0.25 * k)
            joints.append(lib.b2CreateDistanceJoint(world, C.byref(d)))
        # prismatic: limited slider with a spring, motorised slider on a tilted axis
        for i, (spring, motor, axis) in enumerate(((True, False, (1.0, 0.0)), (False, True, (0.8, 0.6)))):
            b = dyn(x0 - 1.0 + 2.0 * i, 3.0)
            d = lib.b2DefaultPrismaticJointDef()
            d.bodyIdA, d.bodyIdB = anchor, b
            d.localAnchorA = A.Vec2(_f32(-1.0 + 2.0 * i), -1.0)
            d.localAnchorB = A.Vec2(0.0, 0.0)
            d.localAxisA = A.Vec2(*axis)
            d.enableSpring, d.hertz, d.dampingRatio, d.targetTranslation = spring, 1.5, 0.2, 0.25
            d.enableLimit, d.lowerTranslation, d.upperTranslation = True, -1.0, 1.25
            d.enableMotor, d.maxMotorForce, d.motorSpeed = motor, 30.0, _f32(1.0 + 0.5 * k)
            joints.append(lib.b2CreatePrismaticJoint(world, C.byref(d)))
        # weld: rigid and soft (both frequencies)
        prev = anchor
        for i, (lh, ah) in enumerate(((0.0, 0.0), (3.0, 2.0), (0.0, 4.0))):
            b = dyn(x0 + 3.0 + 0.7 * i, 4.0)
            d = lib.b2DefaultWeldJointDef()
            d.bodyIdA, d.bodyIdB = prev, b
            d.localAnchorA = A.Vec2(3.0 if i == 0 else 0.35, 0.0)
            d.localAnchorB = A.Vec2(-0.35, 0.0)
            d.linearHertz, d.linearDampingRatio = lh, 0.5
            d.angularHertz, d.angularDampingRatio = ah, 0.7
            joints.append(lib.b2CreateWeldJoint(world, C.byref(d)))
            prev = b
        # wheel: a two-wheeled cart on the ground, spring suspension, one driven wheel, limits
        chassis = dyn(x0, -4.9)
        for i, side in enumerate((-0.6, 0.6)):
            wb = dyn(x0 + side, -5.5, "wheel")
            d = lib.b2DefaultWheelJointDef()
            d.bodyIdA, d.bodyIdB = chassis, wb
            d.localAnchorA = A.Vec2(side, -0.5)
            d.localAnchorB = A.Vec2(0.0, 0.0)
            d.localAxisA = A.Vec2(0.0, 1.0)
            d.enableSpring, d.hertz, d.dampingRatio = True, 4.0, 0.7
            d.enableLimit, d.lowerTranslation, d.upperTranslation = True, -0.25, 0.25
            d.enableMotor, d.maxMotorTorque, d.motorSpeed = i == 0, 5.0, _f32(-3.0 - k)
            joints.append(lib.b2CreateWheelJoint(world, C.byref(d)))
        # motor joint holding a box at an offset from the anchor; mouse joint dragging a loose box; filter joint
        b = dyn(x0 + 6.0, 2.5)
        d = lib.b2DefaultMotorJointDef()
        d.bodyIdA, d.bodyIdB = anchor, b
        d.linearOffset = A.Vec2(6.0, -1.0)
        d.angularOffset = _f32(0.25 * (k + 1))
        d.maxForce, d.maxTorque, d.correctionFactor = 200.0, 20.0, 0.3
        joints.append(lib.b2CreateMotorJoint(world, C.byref(d)))
        b = dyn(x0 + 8.0, 1.0, angle=0.3)
        d = lib.b2DefaultMouseJointDef()
        d.bodyIdA, d.bodyIdB = anchor, b
        d.target = A.Vec2(_f32(x0 + 7.5), 2.0)
        d.hertz, d.dampingRatio, d.maxForce = 5.0, 0.7, 300.0
        joints.append(lib.b2CreateMouseJoint(world, C.byref(d)))
        b1, b2 = dyn(x0 - 8.0, -5.6), dyn(x0 - 8.1, -5.0)
        d = lib.b2DefaultFilterJointDef()
        d.bodyIdA, d.bodyIdB = b1, b2
        joints.append(lib.b2CreateFilterJoint(world, C.byref(d)))
        # revolute with spring + limit + motor for completeness
        b = dyn(x0 + 9.5, 3.5)
        d = lib.b2DefaultRevoluteJointDef()
        d.bodyIdA, d.bodyIdB = anchor, b
        d.localAnchorA = A.Vec2(9.5, 0.0)
        d.localAnchorB = A.Vec2(0.0, 0.5)
        d.enableSpring, d.hertz, d.dampingRatio = True, 1.0, 0.1
        d.enableLimit, d.lowerAngle, d.upperAngle = True, -0.5, 0.75
        d.enableMotor, d.maxMotorTorque, d.motorSpeed = True, 2.0, 1.0
        joints.append(lib.b2CreateRevoluteJoint(world, C.byref(d)))
    scene = Scene(lib, world, bodies, "joint_zoo_%d" % sets)
    scene.joints = joints
    return scene


def sensor_field(lib, count=30, create=None, **world_kw):
    """Sensor shapes (static zones, a kinematic sweeper, a sensor riding on a dynamic body) with boxes, circles and
    capsules raining through them; some visitors opt out of sensor events."""
    world = _world(lib, create=create, **world_kw)
    bodies = [_static_segment(lib, world, (-30.0, 0.0), (30.0, 0.0))]
    sensors = []
    sd = lib.b2DefaultShapeDef()
    ssd = lib.b2DefaultShapeDef()
    ssd.isSensor = True
    ssd.enableSensorEvents = True
    # static zones
    for k, x in enumerate((-8.0, 0.0, 8.0)):
        bd = lib.b2DefaultBodyDef()
        bd.position = A.Vec2(x, 3.0 + k)
        zone = lib.b2CreateBody(world, C.byref(bd))
        bodies.append(zone)
        if k == 1:
            c = A.Circle(A.Vec2(0.0, 0.0), 1.5)
            sensors.append(lib.b2CreateCircleShape(zone, C.byref(ssd), C.byref(c)))
        else:
            box = lib.b2MakeBox(2.0, 0.75)
            sensors.append(lib.b2CreatePolygonShape(zone, C.byref(ssd), C.byref(box)))
    # kinematic sweeper
    bd = lib.b2DefaultBodyDef()
    bd.type = 1
    bd.position = A.Vec2(-12.0, 1.5)
    bd.linearVelocity = A.Vec2(3.0, 0.0)
    sweeper = lib.b2CreateBody(world, C.byref(bd))
    bodies.append(sweeper)
    cap = A.Capsule(A.Vec2(0.0, -1.0), A.Vec2(0.0, 1.0), 0.4)
    sensors.append(lib.b2CreateCapsuleShape(sweeper, C.byref(ssd), C.byref(cap)))
    # a dynamic carrier with a solid box and a larger sensor halo
    bd = lib.b2DefaultBodyDef()
    bd.type = 2
    bd.position = A.Vec2(4.0, 9.0)
    carrier = lib.b2CreateBody(world, C.byref(bd))
    bodies.append(carrier)
    lib.b2CreatePolygonShape(carrier, C.byref(sd), C.byref(_box(lib, 0.4)))
    halo = A.Circle(A.Vec2(0.0, 0.0), 1.6)
    sensors.append(lib.b2CreateCircleShape(carrier, C.byref(ssd), C.byref(halo)))
    state = [12345]

    def rnd():
        state[0] = (1103515245 * state[0] + 12345) & 0x7FFFFFFF
        return state[0] / float(0x7FFFFFFF)

    for i in range(count):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.position = A.Vec2(_f32(-12.0 + 24.0 * rnd()), _f32(6.0 + 10.0 * rnd()))
        b = lib.b2CreateBody(world, C.byref(bd))
        bodies.append(b)
        vsd = lib.b2DefaultShapeDef()
        vsd.enableSensorEvents = (i % 5) != 0
        kind = i % 3
        if kind == 0:
            lib.b2CreatePolygonShape(b, C.byref(vsd), C.byref(_box(lib, 0.3)))
        elif kind == 1:
            c = A.Circle(A.Vec2(0.0, 0.0), 0.3)
            lib.b2CreateCircleShape(b, C.byref(vsd), C.byref(c))
        else:
            c = A.Capsule(A.Vec2(-0.3, 0.0), A.Vec2(0.3, 0.0), 0.2)
            lib.b2CreateCapsuleShape(b, C.byref(vsd), C.byref(c))
    scene = Scene(lib, world, bodies, "sensor_field_%d" % count)
    scene.sensors = sensors
    return scene


def chain_terrain(lib, count=36, create=None, **world_kw):
    """Chain shapes: an open terrain polyline with convex and concave corners and a closed (loop) basin, with boxes,
    circles, capsules and a rounded polygon dropped on them: every chain-segment manifold function, ghost-vertex
    (smooth collision) decisions and the persistent GJK cache."""
    world = _world(lib, create=create, **world_kw)
    bd = lib.b2DefaultBodyDef()
    ground = lib.b2CreateBody(world, C.byref(bd))
    bodies = [ground]
    # open chain: ghost, then the terrain right to left (chain normals point to the right of the travel direction)
    pts = [(22.0, 3.0), (18.0, 0.0), (13.0, 0.0), (10.0, 1.5), (7.0, 0.5), (4.0, 0.5), (2.0, -1.0), (-2.0, -1.0),
           (-4.0, 0.25), (-8.0, 0.0), (-11.0, 2.0), (-14.0, 0.0), (-18.0, 0.0), (-22.0, 4.0)]
    arr = (A.Vec2 * len(pts))(*[A.Vec2(_f32(x), _f32(y)) for x, y in pts])
    cd = lib.b2DefaultChainDef()
    cd.points = arr
    cd.count = len(pts)
    cd.isLoop = False
    chains = [lib.b2CreateChain(ground, C.byref(cd))]
    # closed basin above the terrain (counter-clockwise loop = solid outside, bodies live inside)
    loop = [(-3.0, 6.0), (3.0, 6.0), (4.0, 9.0), (0.0, 8.0), (-4.0, 9.0)]
    mats = (A.SurfaceMaterial * len(loop))()
    for i in range(len(loop)):
        mats[i].friction = _f32(0.2 + 0.15 * i)
        mats[i].restitution = _f32(0.05 * i)
    larr = (A.Vec2 * len(loop))(*[A.Vec2(_f32(x), _f32(y)) for x, y in reversed(loop)])
    cd2 = lib.b2DefaultChainDef()
    cd2.points = larr
    cd2.count = len(loop)
    cd2.materials = mats
    cd2.materialCount = len(loop)
    cd2.isLoop = True
    chains.append(lib.b2CreateChain(ground, C.byref(cd2)))
    state = [2024]

    def rnd():
        state[0] = (1103515245 * state[0] + 12345) & 0x7FFFFFFF
        return state[0] / float(0x7FFFFFFF)

    sd = lib.b2DefaultShapeDef()
    for i in range(count):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        inside = i % 6 == 0
        if inside:
            bd.position = A.Vec2(_f32(-2.0 + 4.0 * rnd()), _f32(6.8 + 0.8 * rnd()))
        else:
            bd.position = A.Vec2(_f32(-19.0 + 38.0 * rnd()), _f32(5.0 + 8.0 * rnd()))
        bd.angularVelocity = _f32(4.0 * rnd() - 2.0)
        b = lib.b2CreateBody(world, C.byref(bd))
        bodies.append(b)
        kind = i % 4
        if kind == 0:
            lib.b2CreatePolygonShape(b, C.byref(sd), C.byref(_box(lib, 0.25 if inside else 0.4)))
        elif kind == 1:
            c = A.Circle(A.Vec2(0.0, 0.0), 0.3)
            lib.b2CreateCircleShape(b, C.byref(sd), C.byref(c))
        elif kind == 2:
            c = A.Capsule(A.Vec2(-0.35, 0.0), A.Vec2(0.35, 0.0), 0.2)
            lib.b2CreateCapsuleShape(b, C.byref(sd), C.byref(c))
        else:
            poly = lib.b2MakeOffsetRoundedBox(0.3, 0.2, A.Vec2(0.0, 0.0), A.Rot(1.0, 0.0), 0.1)
            lib.b2CreatePolygonShape(b, C.byref(sd), C.byref(poly))
    scene = Scene(lib, world, bodies, "chain_terrain_%d" % count)
    scene.chains = chains
    scene._keep = (arr, larr, mats)
    return scene


def random_world(lib, seed=1, count=48, pre_solve_events=False, create=None, **world_kw):
    """Seeded random scene for differential testing: every body type (static, kinematic, dynamic; bullets, fixed
    rotation, damping, gravity scale, sleep thresholds), every shape type with random materials (friction, restitution,
    rolling resistance, tangent speed), random filters (categories, masks, groups), sensors, hit events, multi-shape
    bodies and random revolute / distance / weld joints between neighbours, inside a box of segments."""
    import math
    rng = _Lcg(1000 + 7919 * seed)
    r = rng.next
    world = _world(lib, create=create, enable_continuous=(seed % 5 != 4), **world_kw)
    bd = lib.b2DefaultBodyDef()
    ground = lib.b2CreateBody(world, C.byref(bd))
    gsd = lib.b2DefaultShapeDef()
    gsd.enableHitEvents = True
    chains = []
    if seed % 3 == 1 and not pre_solve_events:
        # every third world stands in a bumpy basin made of ONE chain shape (ghost vertices, chain-segment manifolds with
        # their GJK cache) instead of three segments; travel direction right to left = solid side below
        pts = [(14.0, 32.0), (12.0, 30.0), (10.0, 0.0), (6.0, 0.6), (3.0, -0.4), (0.0, 0.3), (-4.0, -0.5), (-7.0, 0.4),
               (-10.0, 0.0), (-12.0, 30.0), (-14.0, 32.0)]
        arr = (A.Vec2 * len(pts))(*[A.Vec2(_f32(x), _f32(y)) for x, y in pts])
        cd = lib.b2DefaultChainDef()
        cd.points = arr
        cd.count = len(pts)
        cd.isLoop = False
        chains.append(lib.b2CreateChain(ground, C.byref(cd)))
    else:
        for p1, p2 in (((-10.0, 0.0), (10.0, 0.0)), ((-10.0, 0.0), (-12.0, 30.0)), ((10.0, 0.0), (12.0, 30.0))):
            seg = A.Segment(A.Vec2(*p1), A.Vec2(*p2))
            lib.b2CreateSegmentShape(ground, C.byref(gsd), C.byref(seg))
    bodies = [ground]

    def shape_def():
        sd = lib.b2DefaultShapeDef()
        sd.density = _f32(0.5 + 3.0 * r())
        sd.material.friction = _f32(r())
        sd.material.restitution = _f32(0.6 * r()) if r() < 0.4 else 0.0
        sd.material.rollingResistance = _f32(0.2 * r()) if r() < 0.25 else 0.0
        sd.material.tangentSpeed = _f32(2.0 * r() - 1.0) if r() < 0.15 else 0.0
        if r() < 0.3:
            sd.filter.categoryBits = 1 << int(4 * r())
            sd.filter.maskBits = int(r() * 16) | 1
        if r() < 0.2:
            sd.filter.groupIndex = int(5 * r()) - 2
        sd.enableHitEvents = r() < 0.3
        sd.enableContactEvents = r() < 0.8
        sd.enablePreSolveEvents = pre_solve_events and r() < 0.6
        if r() < 0.08:
            sd.isSensor = True
            sd.enableSensorEvents = True
        else:
            sd.enableSensorEvents = r() < 0.7
        return sd

    def add_shape(body, body_type):
        sd = shape_def()
        kind = int(6 * r())
        if pre_solve_events:
            # The reference's continuous pass hands b2PreSolveFcn a temporary manifold computed WITHOUT the shape-order
            # swap of the discrete path (solver.c:366-379 -> contact.c:635-640): a pair whose manifold function is
            # registered the other way round (or not at all: segment vs segment) crashes the reference itself. Keep the
            # static side polygonal and the moving side free of segments so that every such pair is in primary order.
            if body_type == 0:
                kind = 2 + kind % 3
            elif kind == 5:
                kind = 0
        if kind == 0:
            c = A.Circle(A.Vec2(_f32(0.3 * r() - 0.15), 0.0), _f32(0.15 + 0.35 * r()))
            lib.b2CreateCircleShape(body, C.byref(sd), C.byref(c))
        elif kind == 1:
            cap = A.Capsule(A.Vec2(_f32(-0.2 - 0.4 * r()), 0.0), A.Vec2(_f32(0.2 + 0.4 * r()), _f32(0.3 * r())), _f32(0.1 + 0.2 * r()))
            lib.b2CreateCapsuleShape(body, C.byref(sd), C.byref(cap))
        elif kind == 2:
            box = lib.b2MakeOffsetRoundedBox(_f32(0.15 + 0.5 * r()), _f32(0.15 + 0.4 * r()), A.Vec2(_f32(0.2 * r()), 0.0),
                                             A.Rot(1.0, 0.0), 0.0)
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(box))
        elif kind == 3:
            a = r() * 6.28318
            box = lib.b2MakeOffsetRoundedBox(_f32(0.2 + 0.3 * r()), _f32(0.1 + 0.3 * r()), A.Vec2(0.0, 0.0),
                                             A.Rot(_f32(math.cos(a)), _f32(math.sin(a))), _f32(0.02 + 0.1 * r()))
            lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(box))
        elif kind == 4:
            n = 3 + int(6 * r())
            pts = (A.Vec2 * n)()
            for k in range(n):
                ang = 6.28318 * (k + 0.7 * r()) / n
                rad = 0.25 + 0.35 * r()
                pts[k] = A.Vec2(_f32(rad * math.cos(ang)), _f32(rad * math.sin(ang)))
            hull = lib.b2ComputeHull(pts, n)
            if hull.count >= 3:
                poly = lib.b2MakePolygon(C.byref(hull), _f32(0.05 * r()) if r() < 0.3 else 0.0)
                lib.b2CreatePolygonShape(body, C.byref(sd), C.byref(poly))
            else:
                c = A.Circle(A.Vec2(0.0, 0.0), 0.3)
                lib.b2CreateCircleShape(body, C.byref(sd), C.byref(c))
        else:
            seg = A.Segment(A.Vec2(_f32(-0.3 - 0.5 * r()), 0.0), A.Vec2(_f32(0.3 + 0.5 * r()), _f32(0.4 * r() - 0.2)))
            lib.b2CreateSegmentShape(body, C.byref(sd), C.byref(seg))

    dynamic = []
    for i in range(count):
        bd = lib.b2DefaultBodyDef()
        t = r()
        bd.type = 2 if t < 0.82 else (1 if t < 0.9 else 0)
        bd.position = A.Vec2(_f32(-8.0 + 16.0 * r()), _f32(1.0 + 0.6 * i + r()))
        a = 6.28318 * r()
        bd.rotation = A.Rot(_f32(math.cos(a)), _f32(math.sin(a)))
        if bd.type != 0:
            bd.linearVelocity = A.Vec2(_f32(6.0 * r() - 3.0), _f32(4.0 * r() - 8.0 if r() < 0.2 else 2.0 * r() - 1.0))
            bd.angularVelocity = _f32(6.0 * r() - 3.0)
        if bd.type == 1:
            bd.linearVelocity = A.Vec2(_f32(r() - 0.5), _f32(0.4 * r() - 0.2))
        bd.linearDamping = _f32(0.5 * r()) if r() < 0.3 else 0.0
        bd.angularDamping = _f32(0.5 * r()) if r() < 0.3 else 0.0
        bd.gravityScale = _f32(0.5 + r()) if r() < 0.2 else 1.0
        bd.fixedRotation = r() < 0.1
        bd.isBullet = bd.type == 2 and r() < 0.12 and not pre_solve_events  # bullets sweep against shapes of every type
        bd.allowFastRotation = r() < 0.1
        bd.enableSleep = r() < 0.9
        bd.sleepThreshold = _f32(0.05 + 0.2 * r()) if r() < 0.2 else _f32(0.05)
        body = lib.b2CreateBody(world, C.byref(bd))
        for _ in range(1 + (1 if r() < 0.25 else 0) + (1 if r() < 0.1 else 0)):
            add_shape(body, bd.type)
        bodies.append(body)
        if bd.type == 2:
            dynamic.append((body, bd.position.x, bd.position.y, A.Rot(bd.rotation.c, bd.rotation.s)))
    # a few fast bullets aimed at the pile
    for i in range(0 if pre_solve_events else 3):
        bd = lib.b2DefaultBodyDef()
        bd.type = 2
        bd.isBullet = True
        bd.position = A.Vec2(_f32(-9.0 + 0.5 * i), _f32(20.0 + 3.0 * i))
        bd.linearVelocity = A.Vec2(_f32(40.0 + 20.0 * r()), _f32(-60.0 * r()))
        body = lib.b2CreateBody(world, C.byref(bd))
        sd = lib.b2DefaultShapeDef()
        sd.enableHitEvents = True
        c = A.Circle(A.Vec2(0.0, 0.0), 0.12)
        lib.b2CreateCircleShape(body, C.byref(sd), C.byref(c))
        bodies.append(body)
    joints = []

    def local(rot, dx, dy):     # world-space offset -> the body's local frame
        return A.Vec2(_f32(rot.c * dx + rot.s * dy), _f32(-rot.s * dx + rot.c * dy))

    # joints only between vertical neighbours (consistent anchors at the midpoint, no loops): a random scene must not
    # blow up, the point is to reach the rarely taken branches, not NaN arithmetic
    for k in range(0, len(dynamic) - 1, 2):
        if r() < 0.45:
            continue
        (a, ax, ay, qa), (b, bx, by, qb) = dynamic[k], dynamic[k + 1]
        mx, my = 0.5 * (ax + bx), 0.5 * (ay + by)
        kind = int(3 * r())     # prismatic / wheel joints between tumbling bodies blow up in the reference itself: joint_zoo has them
        if kind == 0:
            d = lib.b2DefaultRevoluteJointDef()
            d.bodyIdA, d.bodyIdB = a, b
            d.localAnchorA, d.localAnchorB = local(qa, mx - ax, my - ay), local(qb, mx - bx, my - by)
            d.enableLimit, d.lowerAngle, d.upperAngle = r() < 0.5, -0.5, 0.8
            d.enableMotor, d.maxMotorTorque, d.motorSpeed = r() < 0.3, 5.0, _f32(2.0 * r() - 1.0)
            d.enableSpring, d.hertz, d.dampingRatio = r() < 0.3, 2.0, 0.3
            d.collideConnected = r() < 0.3
            joints.append(lib.b2CreateRevoluteJoint(world, C.byref(d)))
        elif kind == 1:
            d = lib.b2DefaultDistanceJointDef()
            d.bodyIdA, d.bodyIdB = a, b
            d.length = _f32(max(0.2, math.hypot(bx - ax, by - ay)))
            d.enableSpring, d.hertz, d.dampingRatio = r() < 0.6, _f32(1.0 + 4.0 * r()), _f32(r())
            d.enableLimit, d.minLength, d.maxLength = r() < 0.4, _f32(0.5 * d.length), _f32(1.5 * d.length)
            joints.append(lib.b2CreateDistanceJoint(world, C.byref(d)))
        else:
            d = lib.b2DefaultWeldJointDef()
            d.bodyIdA, d.bodyIdB = a, b
            d.localAnchorA, d.localAnchorB = local(qa, mx - ax, my - ay), local(qb, mx - bx, my - by)
            d.referenceAngle = _f32(math.atan2(qb.s, qb.c) - math.atan2(qa.s, qa.c))
            d.linearHertz, d.angularHertz = _f32(3.0 * r()), _f32(3.0 * r())
            joints.append(lib.b2CreateWeldJoint(world, C.byref(d)))
    scene = Scene(lib, world, bodies, "random_world_%d" % seed)
    scene.joints = joints
    scene.chains = chains
    return scene


SCENES = {
    "chain_terrain": chain_terrain,
    "sensor_field": sensor_field,
    "joint_zoo": joint_zoo,
    "random_world": random_world,
    "bench2d": bench2d,
    "large_pyramid": large_pyramid,
    "many_pyramids": many_pyramids,
    "joint_grid": joint_grid,
    "falling_shapes": falling_shapes,
    "polygon_soup": polygon_soup,
    "jointed_piles": jointed_piles,
}
