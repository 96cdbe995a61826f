"""Python mirror of forge2d's Dart API for the ``World.step`` path (host side above the C ABI).

Dart is not available in the build image, so the idiomatic layer that sits above ``RawBox2D`` in the reference
(``packages/forge2d/lib/src/api/*.dart``) is mirrored here with the same class names, member names, argument meaning
and error behaviour, over ctypes instead of ``dart:ffi``:

    World            packages/forge2d/lib/src/api/world.dart:27-120 (step, gravity, toggles), :386-510 (events)
    Body / BodyDef   packages/forge2d/lib/src/api/body.dart, defs.dart
    Shape / ShapeDef packages/forge2d/lib/src/api/shape.dart, defs.dart
    Polygon.box/square -> b2MakeOffsetRoundedBox   packages/forge2d/lib/src/backend/raw_box2d_ffi.dart:682-690

Like the Dart layer it only talks to the backend through primitive calls, keeps user data on this side of the boundary,
tracks ``locked`` while stepping (``StateError`` -> ``RuntimeError``) and copies events out right after the call.
Every ``World.step`` runs on the GPU through ``b2World_Step`` of ``libforge2d_b200.so``.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Any, List, Optional

from . import _abi as A


class StateError(RuntimeError):
    """Dart's StateError: a mutating call was made while the world is stepping."""


@dataclass(frozen=True)
class Vector2:
    x: float = 0.0
    y: float = 0.0

    def __iter__(self):
        return iter((self.x, self.y))


@dataclass(frozen=True)
class Rot:
    c: float = 1.0
    s: float = 0.0

    @staticmethod
    def fromAngle(angle: float) -> "Rot":  # math.dart:20 (dart:math cos/sin in double, passed through as floats)
        return Rot(math.cos(angle), math.sin(angle))

    @property
    def angle(self) -> float:
        return math.atan2(self.s, self.c)


@dataclass(frozen=True)
class Transform:
    p: Vector2 = Vector2()
    q: Rot = Rot()


class BodyType(IntEnum):
    static = 0
    kinematic = 1
    dynamic = 2


@dataclass
class WorldDef:  # defs.dart WorldDef -> the 10 flattened createWorld parameters (world.dart:37-49)
    gravity: Vector2 = Vector2(0.0, -10.0)
    restitutionThreshold: float = 1.0
    hitEventThreshold: float = 1.0
    contactHertz: float = 30.0
    contactDampingRatio: float = 10.0
    maxContactPushSpeed: float = 3.0
    maximumLinearSpeed: float = 400.0
    enableSleep: bool = True
    enableContinuous: bool = True


@dataclass
class BodyDef:
    type: BodyType = BodyType.static
    position: Vector2 = Vector2()
    rotation: Rot = Rot()
    linearVelocity: Vector2 = Vector2()
    angularVelocity: float = 0.0
    linearDamping: float = 0.0
    angularDamping: float = 0.0
    gravityScale: float = 1.0
    sleepThreshold: float = 0.05
    enableSleep: bool = True
    isAwake: bool = True
    fixedRotation: bool = False
    isBullet: bool = False
    isEnabled: bool = True
    allowFastRotation: bool = False
    name: Optional[str] = None
    userData: Any = None


@dataclass
class ShapeDef:
    friction: float = 0.6
    restitution: float = 0.0
    rollingResistance: float = 0.0
    tangentSpeed: float = 0.0
    density: float = 1.0
    categoryBits: int = 1
    maskBits: int = 0xFFFFFFFFFFFFFFFF
    groupIndex: int = 0
    isSensor: bool = False
    enableSensorEvents: bool = False
    enableContactEvents: bool = True
    enableHitEvents: bool = False
    enablePreSolveEvents: bool = False
    userData: Any = None


@dataclass(frozen=True)
class Circle:
    center: Vector2 = Vector2()
    radius: float = 0.5


@dataclass(frozen=True)
class Capsule:
    center1: Vector2 = Vector2(-0.5, 0.0)
    center2: Vector2 = Vector2(0.5, 0.0)
    radius: float = 0.25


@dataclass(frozen=True)
class Segment:
    point1: Vector2 = Vector2()
    point2: Vector2 = Vector2(1.0, 0.0)


@dataclass(frozen=True)
class Polygon:
    """Box-like polygons; geometry.dart Polygon.box / Polygon.square go through b2MakeOffsetRoundedBox."""
    halfWidth: float
    halfHeight: float
    center: Vector2 = Vector2()
    rotation: Rot = Rot()
    radius: float = 0.0
    points: Optional[tuple] = None  # arbitrary convex point list -> b2ComputeHull + b2MakePolygon

    @staticmethod
    def box(halfWidth, halfHeight, center=Vector2(), rotation=Rot(), radius=0.0):
        return Polygon(halfWidth, halfHeight, center, rotation, radius)

    @staticmethod
    def square(halfExtent):
        return Polygon(halfExtent, halfExtent)

    @staticmethod
    def fromPoints(points, radius=0.0):
        return Polygon(0.0, 0.0, radius=radius, points=tuple(points))


@dataclass
class RevoluteJointDef:
    bodyA: "Body" = None
    bodyB: "Body" = None
    localAnchorA: Vector2 = Vector2()
    localAnchorB: Vector2 = Vector2()
    referenceAngle: float = 0.0
    enableSpring: bool = False
    hertz: float = 0.0
    dampingRatio: float = 0.0
    enableLimit: bool = False
    lowerAngle: float = 0.0
    upperAngle: float = 0.0
    enableMotor: bool = False
    maxMotorTorque: float = 0.0
    motorSpeed: float = 0.0
    collideConnected: bool = False
    userData: Any = None


# ---- events (events.dart) -----------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ContactPoint:
    point: Vector2
    separation: float


@dataclass(frozen=True)
class ContactBeginEvent:
    shapeA: "Shape"
    shapeB: "Shape"
    normal: Vector2
    points: List[ContactPoint]


@dataclass(frozen=True)
class ContactEndEvent:
    shapeA: "Shape"
    shapeB: "Shape"


@dataclass(frozen=True)
class ContactHitEvent:
    shapeA: "Shape"
    shapeB: "Shape"
    point: Vector2
    normal: Vector2
    approachSpeed: float


@dataclass(frozen=True)
class ContactEvents:
    begin: List[ContactBeginEvent] = field(default_factory=list)
    end: List[ContactEndEvent] = field(default_factory=list)
    hit: List[ContactHitEvent] = field(default_factory=list)


@dataclass(frozen=True)
class SensorEvent:
    sensor: "Shape"
    visitor: "Shape"


@dataclass(frozen=True)
class SensorEvents:
    begin: List[SensorEvent] = field(default_factory=list)
    end: List[SensorEvent] = field(default_factory=list)


@dataclass(frozen=True)
class BodyMoveEvent:
    body: "Body"
    transform: Transform
    fellAsleep: bool


_backend = None


def initializeForge2D(library=None):
    """initialize.dart:3-8 — selects the backend. Default: the CUDA library (raises if it is not built)."""
    global _backend
    if library is None:
        from . import load_library
        library = load_library()
    _backend = library
    return library


def _lib():
    if _backend is None:
        initializeForge2D()
    return _backend


def _v(v):
    return A.Vec2(float(v.x), float(v.y))


class Shape:
    def __init__(self, world, raw):
        self.world, self._id = world, raw

    @staticmethod
    def internal(world, raw):
        return Shape(world, raw)

    @property
    def key(self):
        return (self._id.index1, self._id.generation)

    def __eq__(self, other):
        return isinstance(other, Shape) and other.world is self.world and other.key == self.key

    def __hash__(self):
        return hash(self.key)

    @property
    def isValid(self):
        return bool(_lib().b2Shape_IsValid(self._id))

    @property
    def body(self):
        return Body(self.world, _lib().b2Shape_GetBody(self._id))

    @property
    def aabb(self):
        a = _lib().b2Shape_GetAABB(self._id)
        return (Vector2(a.lowerBound.x, a.lowerBound.y), Vector2(a.upperBound.x, a.upperBound.y))

    @property
    def userData(self):
        return self.world.shapeUserData.get(self.key)


class Body:
    def __init__(self, world, raw):
        self.world, self._id = world, raw

    @staticmethod
    def internal(world, raw):
        return Body(world, raw)

    @property
    def key(self):
        return (self._id.index1, self._id.generation)

    def __eq__(self, other):
        return isinstance(other, Body) and other.world is self.world and other.key == self.key

    def __hash__(self):
        return hash(self.key)

    @property
    def isValid(self):
        return bool(_lib().b2Body_IsValid(self._id))

    @property
    def type(self):
        return BodyType(_lib().b2Body_GetType(self._id))

    @property
    def position(self):
        p = _lib().b2Body_GetPosition(self._id)
        return Vector2(p.x, p.y)

    @property
    def rotation(self):
        q = _lib().b2Body_GetRotation(self._id)
        return Rot(q.c, q.s)

    @property
    def transform(self):
        return Transform(self.position, self.rotation)

    @property
    def linearVelocity(self):
        v = _lib().b2Body_GetLinearVelocity(self._id)
        return Vector2(v.x, v.y)

    @linearVelocity.setter
    def linearVelocity(self, value):
        self.world.checkCanMutate("set a velocity")
        _lib().b2Body_SetLinearVelocity(self._id, _v(value))

    @property
    def angularVelocity(self):
        return float(_lib().b2Body_GetAngularVelocity(self._id))

    @angularVelocity.setter
    def angularVelocity(self, value):
        self.world.checkCanMutate("set a velocity")
        _lib().b2Body_SetAngularVelocity(self._id, float(value))

    @property
    def mass(self):
        return float(_lib().b2Body_GetMass(self._id))

    @property
    def rotationalInertia(self):
        return float(_lib().b2Body_GetRotationalInertia(self._id))

    @property
    def worldCenterOfMass(self):
        c = _lib().b2Body_GetWorldCenterOfMass(self._id)
        return Vector2(c.x, c.y)

    @property
    def isAwake(self):
        return bool(_lib().b2Body_IsAwake(self._id))

    @property
    def shapeCount(self):
        return int(_lib().b2Body_GetShapeCount(self._id))

    @property
    def userData(self):
        return self.world.bodyUserData.get(self.key)

    def createShape(self, geometry, definition: Optional[ShapeDef] = None) -> Shape:
        """body.dart createShape: dispatches on the geometry type to b2Create{Circle,Capsule,Segment,Polygon}Shape."""
        self.world.checkCanMutate("create a shape")
        lib = _lib()
        d = definition or ShapeDef()
        sd = lib.b2DefaultShapeDef()
        sd.material.friction = d.friction
        sd.material.restitution = d.restitution
        sd.material.rollingResistance = d.rollingResistance
        sd.material.tangentSpeed = d.tangentSpeed
        sd.density = d.density
        sd.filter.categoryBits = d.categoryBits
        sd.filter.maskBits = d.maskBits
        sd.filter.groupIndex = d.groupIndex
        sd.isSensor = d.isSensor
        sd.enableSensorEvents = d.enableSensorEvents
        sd.enableContactEvents = d.enableContactEvents
        sd.enableHitEvents = d.enableHitEvents
        sd.enablePreSolveEvents = d.enablePreSolveEvents
        if isinstance(geometry, Circle):
            g = A.Circle(_v(geometry.center), geometry.radius)
            raw = lib.b2CreateCircleShape(self._id, C.byref(sd), C.byref(g))
        elif isinstance(geometry, Capsule):
            g = A.Capsule(_v(geometry.center1), _v(geometry.center2), geometry.radius)
            raw = lib.b2CreateCapsuleShape(self._id, C.byref(sd), C.byref(g))
        elif isinstance(geometry, Segment):
            g = A.Segment(_v(geometry.point1), _v(geometry.point2))
            raw = lib.b2CreateSegmentShape(self._id, C.byref(sd), C.byref(g))
        elif isinstance(geometry, Polygon):
            if geometry.points is not None:
                pts = (A.Vec2 * len(geometry.points))(*[_v(p) for p in geometry.points])
                hull = lib.b2ComputeHull(pts, len(geometry.points))
                if hull.count == 0:
                    raise ValueError("Polygon.fromPoints: the points do not form a convex hull")
                g = lib.b2MakePolygon(C.byref(hull), geometry.radius)
            else:
                g = lib.b2MakeOffsetRoundedBox(geometry.halfWidth, geometry.halfHeight, _v(geometry.center),
                                               A.Rot(geometry.rotation.c, geometry.rotation.s), geometry.radius)
            raw = lib.b2CreatePolygonShape(self._id, C.byref(sd), C.byref(g))
        else:
            raise TypeError("unsupported geometry %r" % (geometry,))
        shape = Shape(self.world, raw)
        if d.userData is not None:
            self.world.shapeUserData[shape.key] = d.userData
        return shape


class World:
    """world.dart:27 — wraps a native b2WorldId; must only be used from the thread that created it."""

    def __init__(self, gravity: Optional[Vector2] = None, definition: Optional[WorldDef] = None):
        lib = _lib()
        d = definition or WorldDef()
        wd = lib.b2DefaultWorldDef()
        wd.gravity = _v(gravity or d.gravity)
        wd.restitutionThreshold = d.restitutionThreshold
        wd.hitEventThreshold = d.hitEventThreshold
        wd.contactHertz = d.contactHertz
        wd.contactDampingRatio = d.contactDampingRatio
        wd.maxContactPushSpeed = d.maxContactPushSpeed
        wd.maximumLinearSpeed = d.maximumLinearSpeed
        wd.enableSleep = d.enableSleep
        wd.enableContinuous = d.enableContinuous
        self.id = lib.b2CreateWorld(C.byref(wd))
        self.bodyUserData, self.shapeUserData, self.jointUserData = {}, {}, {}
        self.locked = False
        self.deferredActions = []

    def checkCanMutate(self, operation):
        if self.locked:
            raise StateError("Cannot %s while the world is stepping. Collision callbacks run inside step(); create "
                             "after step() returns instead. Destroy operations are deferred automatically." % operation)

    @property
    def isValid(self):
        return bool(_lib().b2World_IsValid(self.id))

    def destroy(self):
        if self.locked:
            self.deferredActions.append(self.destroy)
            return
        _lib().b2DestroyWorld(self.id)

    def step(self, timeStep: float, subStepCount: int = 4):
        """world.dart:105-120 -> RawBox2DFfi.worldStep -> b2World_Step: one synchronous call; on this backend the
        whole step runs as CUDA kernels."""
        assert self.isValid, "World has been destroyed"
        self.locked = True
        try:
            _lib().b2World_Step(self.id, timeStep, subStepCount)
        finally:
            self.locked = False
            if self.deferredActions:
                actions, self.deferredActions = list(self.deferredActions), []
                for action in actions:
                    action()

    @property
    def gravity(self):
        g = _lib().b2World_GetGravity(self.id)
        return Vector2(g.x, g.y)

    @gravity.setter
    def gravity(self, value):
        _lib().b2World_SetGravity(self.id, _v(value))

    @property
    def sleepingEnabled(self):
        return bool(_lib().b2World_IsSleepingEnabled(self.id))

    @sleepingEnabled.setter
    def sleepingEnabled(self, flag):
        _lib().b2World_EnableSleeping(self.id, bool(flag))

    @property
    def continuousEnabled(self):
        return bool(_lib().b2World_IsContinuousEnabled(self.id))

    @continuousEnabled.setter
    def continuousEnabled(self, flag):
        _lib().b2World_EnableContinuous(self.id, bool(flag))

    @property
    def awakeBodyCount(self):
        return int(_lib().b2World_GetAwakeBodyCount(self.id))

    def createBody(self, definition: Optional[BodyDef] = None) -> Body:
        self.checkCanMutate("create a body")
        lib = _lib()
        d = definition or BodyDef()
        bd = lib.b2DefaultBodyDef()
        bd.type = int(d.type)
        bd.position = _v(d.position)
        bd.rotation = A.Rot(d.rotation.c, d.rotation.s)
        bd.linearVelocity = _v(d.linearVelocity)
        bd.angularVelocity = d.angularVelocity
        bd.linearDamping = d.linearDamping
        bd.angularDamping = d.angularDamping
        bd.gravityScale = d.gravityScale
        bd.sleepThreshold = d.sleepThreshold
        bd.enableSleep = d.enableSleep
        bd.isAwake = d.isAwake
        bd.fixedRotation = d.fixedRotation
        bd.isBullet = d.isBullet
        if d.name:
            bd.name = d.name.encode()
        bd.isEnabled = d.isEnabled
        bd.allowFastRotation = d.allowFastRotation
        body = Body(self, lib.b2CreateBody(self.id, C.byref(bd)))
        if d.userData is not None:
            self.bodyUserData[body.key] = d.userData
        return body

    def createRevoluteJoint(self, definition: RevoluteJointDef):
        self.checkCanMutate("create a joint")
        lib = _lib()
        d = definition
        jd = lib.b2DefaultRevoluteJointDef()
        jd.bodyIdA, jd.bodyIdB = d.bodyA._id, d.bodyB._id
        jd.localAnchorA, jd.localAnchorB = _v(d.localAnchorA), _v(d.localAnchorB)
        jd.referenceAngle = d.referenceAngle
        jd.enableSpring, jd.hertz, jd.dampingRatio = d.enableSpring, d.hertz, d.dampingRatio
        jd.enableLimit, jd.lowerAngle, jd.upperAngle = d.enableLimit, d.lowerAngle, d.upperAngle
        jd.enableMotor, jd.maxMotorTorque, jd.motorSpeed = d.enableMotor, d.maxMotorTorque, d.motorSpeed
        jd.collideConnected = d.collideConnected
        return lib.b2CreateRevoluteJoint(self.id, C.byref(jd))

    # ---- events: poll after each step; the returned collections are copies (world.dart:378-510)
    @property
    def contactEvents(self) -> ContactEvents:
        ev = _lib().b2World_GetContactEvents(self.id)
        begin = []
        for i in range(ev.beginCount):
            e = ev.beginEvents[i]
            m = e.manifold
            begin.append(ContactBeginEvent(Shape(self, _copy(e.shapeIdA)), Shape(self, _copy(e.shapeIdB)),
                                           Vector2(m.normal.x, m.normal.y),
                                           [ContactPoint(Vector2(m.points[k].point.x, m.points[k].point.y),
                                                         m.points[k].separation) for k in range(m.pointCount)]))
        end = [ContactEndEvent(Shape(self, _copy(ev.endEvents[i].shapeIdA)), Shape(self, _copy(ev.endEvents[i].shapeIdB)))
               for i in range(ev.endCount)]
        hit = []
        for i in range(ev.hitCount):
            e = ev.hitEvents[i]
            hit.append(ContactHitEvent(Shape(self, _copy(e.shapeIdA)), Shape(self, _copy(e.shapeIdB)),
                                       Vector2(e.point.x, e.point.y), Vector2(e.normal.x, e.normal.y), e.approachSpeed))
        return ContactEvents(begin, end, hit)

    # ---- callbacks inside the step (world.dart:616-650): run on the calling thread during step()
    @property
    def customFilterCallback(self):
        return getattr(self, "_customFilter", None)

    @customFilterCallback.setter
    def customFilterCallback(self, callback):
        """bool callback(Shape shapeA, Shape shapeB), consulted for pairs that pass the regular filtering; None clears."""
        lib = _lib()
        self._customFilter = callback
        if callback is None:
            self._customFilterNative = None
            lib.b2World_SetCustomFilterCallback(self.id, None, None)
            return
        fcn = C.CFUNCTYPE(C.c_bool, A.ShapeId, A.ShapeId, C.c_void_p)(
            lambda a, b, ctx: bool(callback(Shape.internal(self, _copy(a)), Shape.internal(self, _copy(b)))))
        self._customFilterNative = fcn  # keeps the trampoline alive (NativeCallable.isolateLocal in the Dart layer)
        lib.b2World_SetCustomFilterCallback(self.id, C.cast(fcn, C.c_void_p), None)

    @property
    def preSolveCallback(self):
        return getattr(self, "_preSolve", None)

    @preSolveCallback.setter
    def preSolveCallback(self, callback):
        """bool callback(Shape shapeA, Shape shapeB, Vector2 normal); returning False disables the contact for the step.
        Only shapes created with enablePreSolveEvents participate. None clears."""
        lib = _lib()
        self._preSolve = callback
        if callback is None:
            self._preSolveNative = None
            lib.b2World_SetPreSolveCallback(self.id, None, None)
            return
        fcn = C.CFUNCTYPE(C.c_bool, A.ShapeId, A.ShapeId, C.POINTER(A.Manifold), C.c_void_p)(
            lambda a, b, m, ctx: bool(callback(Shape.internal(self, _copy(a)), Shape.internal(self, _copy(b)),
                                               Vector2(m.contents.normal.x, m.contents.normal.y))))
        self._preSolveNative = fcn
        lib.b2World_SetPreSolveCallback(self.id, C.cast(fcn, C.c_void_p), None)

    def draw(self, debugDraw: "DebugDraw"):
        """world.dart:664-760 -> b2World_Draw: emits the world's debug geometry into `debugDraw`."""
        lib = _lib()
        v = lambda p: Vector2(p.x, p.y)
        t = lambda x: Transform(Vector2(x.p.x, x.p.y), Rot(x.q.c, x.q.s))
        native = lib.b2DefaultDebugDraw()
        keep = [
            A.DrawPolygonFcn(lambda vs, n, color, ctx: debugDraw.drawPolygon([v(vs[i]) for i in range(n)], color)),
            A.DrawSolidPolygonFcn(lambda xf, vs, n, r, color, ctx: debugDraw.drawSolidPolygon(t(xf), [v(vs[i]) for i in range(n)], r, color)),
            A.DrawCircleFcn(lambda c, r, color, ctx: debugDraw.drawCircle(v(c), r, color)),
            A.DrawSolidCircleFcn(lambda xf, r, color, ctx: debugDraw.drawSolidCircle(t(xf), r, color)),
            A.DrawSolidCapsuleFcn(lambda p1, p2, r, color, ctx: debugDraw.drawSolidCapsule(v(p1), v(p2), r, color)),
            A.DrawSegmentFcn(lambda p1, p2, color, ctx: debugDraw.drawSegment(v(p1), v(p2), color)),
            A.DrawTransformFcn(lambda xf, ctx: debugDraw.drawTransform(t(xf))),
            A.DrawPointFcn(lambda p, size, color, ctx: debugDraw.drawPoint(v(p), size, color)),
            A.DrawStringFcn(lambda p, text, color, ctx: debugDraw.drawString(v(p), bytes(text).decode(), color)),
        ]
        (native.DrawPolygonFcn, native.DrawSolidPolygonFcn, native.DrawCircleFcn, native.DrawSolidCircleFcn,
         native.DrawSolidCapsuleFcn, native.DrawSegmentFcn, native.DrawTransformFcn, native.DrawPointFcn, native.DrawStringFcn) = keep
        for name in A.DEBUG_DRAW_OPTIONS[1:]:
            setattr(native, name, bool(getattr(debugDraw, name)))
        bounds = debugDraw.drawingBounds
        native.useDrawingBounds = bounds is not None
        if bounds is not None:
            (lx, ly), (ux, uy) = bounds
            native.drawingBounds = A.AABB(A.Vec2(lx, ly), A.Vec2(ux, uy))
        lib.b2World_Draw(self.id, C.byref(native))

    @property
    def sensorEvents(self) -> SensorEvents:
        ev = _lib().b2World_GetSensorEvents(self.id)
        begin = [SensorEvent(Shape(self, _copy(ev.beginEvents[i].sensorShapeId)), Shape(self, _copy(ev.beginEvents[i].visitorShapeId)))
                 for i in range(ev.beginCount)]
        end = [SensorEvent(Shape(self, _copy(ev.endEvents[i].sensorShapeId)), Shape(self, _copy(ev.endEvents[i].visitorShapeId)))
               for i in range(ev.endCount)]
        return SensorEvents(begin, end)

    @property
    def bodyMoveEvents(self) -> List[BodyMoveEvent]:
        ev = _lib().b2World_GetBodyEvents(self.id)
        out = []
        for i in range(ev.moveCount):
            m = ev.moveEvents[i]
            out.append(BodyMoveEvent(Body(self, _copy(m.bodyId)),
                                     Transform(Vector2(m.transform.p.x, m.transform.p.y), Rot(m.transform.q.c, m.transform.q.s)),
                                     bool(m.fellAsleep)))
        return out


def _copy(raw_id):
    return type(raw_id)(raw_id.index1, raw_id.world0, raw_id.generation)


class DebugDraw:
    """debug_draw.dart: receives the world's debug geometry when passed to World.draw. Override the draw methods for
    the primitives you care about and flip the options; `drawingBounds` = ((lx, ly), (ux, uy)) restricts the drawing."""
    drawingBounds = None
    drawShapes = True
    drawJoints = True
    drawJointExtras = False
    drawBounds = False
    drawMass = False
    drawBodyNames = False
    drawContacts = False
    drawGraphColors = False
    drawContactNormals = False
    drawContactImpulses = False
    drawContactFeatures = False
    drawFrictionImpulses = False
    drawIslands = False

    def drawPolygon(self, vertices, color): pass
    def drawSolidPolygon(self, transform, vertices, radius, color): pass
    def drawCircle(self, center, radius, color): pass
    def drawSolidCircle(self, transform, radius, color): pass
    def drawSolidCapsule(self, p1, p2, radius, color): pass
    def drawSegment(self, p1, p2, color): pass
    def drawTransform(self, transform): pass
    def drawPoint(self, p, size, color): pass
    def drawString(self, p, text, color): pass


class WorldBatch:
    """Extension (not in the reference): N device-resident replicas of a template world stepped by one kernel launch per
    step, one thread block per world (f2dBatch_* in include/forge2d_b200.h)."""

    def __init__(self, template, count: int = 0):
        """`template`: one World replicated `count` times, or a list of (different) Worlds, one image each."""
        self._lib = _lib()
        if isinstance(template, (list, tuple)):
            ids = (A.WorldId * len(template))(*[w.id for w in template])
            self._batch = self._lib.f2dBatch_CreateFromWorlds(ids, len(template))
            count = len(template)
        else:
            self._batch = self._lib.f2dBatch_Create(template.id, count)
        if not self._batch:
            raise RuntimeError("f2dBatch_Create failed: %s" % self._lib.f2dGetLastError().decode())
        self.count = count

    def step(self, timeStep: float, subStepCount: int = 4, steps: int = 1):
        """Steps every world; raises when a world ends the call with an error flag up (a world that merely needs more
        contact room is not an error: the batch grows and that world repeats the step)."""
        self._lib.f2dBatch_StepN(self._batch, timeStep, subStepCount, steps)
        self._lib.f2dBatch_Synchronize(self._batch)
        flags = self.errorFlags
        if flags & ~0x20:  # everything but the per-step "an event was dropped" flag
            bad = [i for i, f in enumerate(self.worldErrors()) if f & ~0x20]
            raise RuntimeError("WorldBatch.step: error flags 0x%x in worlds %r" % (flags, bad[:16]))

    def worldErrors(self):
        out = (C.c_uint32 * self.count)()
        self._lib.f2dBatch_GetWorldErrors(self._batch, out, self.count)
        return list(out)

    def bodyTransforms(self, maxBodiesPerWorld: int):
        """numpy view (count, maxBodiesPerWorld) of b2BodyMoveEvent records in pinned memory, plus per-world counts."""
        import numpy as np
        ev, cnt = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
        self._lib.f2dBatch_ReadBodyEvents(self._batch, maxBodiesPerWorld, C.byref(ev), C.byref(cnt))
        return self._view(ev, cnt, maxBodiesPerWorld)

    def _view(self, ev, cnt, maxBodiesPerWorld):
        import numpy as np
        dtype = np.dtype([("p", "<f4", 2), ("q", "<f4", 2), ("bodyIndex1", "<i4"), ("world0", "<u2"), ("generation", "<u2"),
                          ("userData", "<u8"), ("fellAsleep", "u1"), ("pad", "u1", 7)])
        assert dtype.itemsize == C.sizeof(A.BodyMoveEvent)
        n = self.count * maxBodiesPerWorld
        records = np.ctypeslib.as_array(C.cast(ev, C.POINTER(C.c_uint8)), shape=(n * dtype.itemsize,)).view(dtype)
        counts = np.ctypeslib.as_array(cnt, shape=(self.count,))
        return records.reshape(self.count, maxBodiesPerWorld), counts

    def stepAndReadBodyTransforms(self, timeStep: float, maxBodiesPerWorld: int, subStepCount: int = 4):
        """One step of every world and that step's body transforms on the host in one call: the worlds are stepped in
        slices and each slice's read-back overlaps the stepping of the next (f2dBatch_StepAndReadBodyEvents)."""
        ev, cnt = C.POINTER(A.BodyMoveEvent)(), C.POINTER(C.c_int)()
        self._lib.f2dBatch_StepAndReadBodyEvents(self._batch, timeStep, subStepCount, maxBodiesPerWorld, C.byref(ev), C.byref(cnt))
        return self._view(ev, cnt, maxBodiesPerWorld)

    @property
    def errorFlags(self):
        return int(self._lib.f2dBatch_GetErrorFlags(self._batch))

    def destroy(self):
        if self._batch:
            self._lib.f2dBatch_Destroy(self._batch)
            self._batch = None
