"""ctypes mirror of the C ABI in include/forge2d_b200.h (== the Box2D v3.1.1 public API subset forge2d binds through
packages/forge2d/lib/src/ffi/box2d.g.dart).  The same binding code can wrap

* the product library ``forge2d_b200/csrc/libforge2d_b200.so`` (CUDA step), and
* any other library exporting the same ``b2*`` symbols (tests wrap the compiled reference ``oracle/_ref`` and the
  host emulation build with it, so parity tests make *identical* call sequences against both).
"""
import ctypes as C
import os

c_float, c_int, c_bool, c_void_p = C.c_float, C.c_int, C.c_bool, C.c_void_p


class Vec2(C.Structure):
    _fields_ = [("x", c_float), ("y", c_float)]


class Rot(C.Structure):
    _fields_ = [("c", c_float), ("s", c_float)]


class Transform(C.Structure):
    _fields_ = [("p", Vec2), ("q", Rot)]


class AABB(C.Structure):
    _fields_ = [("lowerBound", Vec2), ("upperBound", Vec2)]


class WorldId(C.Structure):
    _fields_ = [("index1", C.c_uint16), ("generation", C.c_uint16)]


class BodyId(C.Structure):
    _fields_ = [("index1", C.c_int32), ("world0", C.c_uint16), ("generation", C.c_uint16)]


class ShapeId(C.Structure):
    _fields_ = [("index1", C.c_int32), ("world0", C.c_uint16), ("generation", C.c_uint16)]


class JointId(C.Structure):
    _fields_ = [("index1", C.c_int32), ("world0", C.c_uint16), ("generation", C.c_uint16)]


class Circle(C.Structure):
    _fields_ = [("center", Vec2), ("radius", c_float)]


class Capsule(C.Structure):
    _fields_ = [("center1", Vec2), ("center2", Vec2), ("radius", c_float)]


class Polygon(C.Structure):
    _fields_ = [("vertices", Vec2 * 8), ("normals", Vec2 * 8), ("centroid", Vec2), ("radius", c_float), ("count", c_int)]


class Segment(C.Structure):
    _fields_ = [("point1", Vec2), ("point2", Vec2)]


class Hull(C.Structure):
    _fields_ = [("points", Vec2 * 8), ("count", c_int)]


class ManifoldPoint(C.Structure):
    _fields_ = [("point", Vec2), ("anchorA", Vec2), ("anchorB", Vec2), ("separation", c_float),
                ("normalImpulse", c_float), ("tangentImpulse", c_float), ("totalNormalImpulse", c_float),
                ("normalVelocity", c_float), ("id", C.c_uint16), ("persisted", c_bool)]


class Manifold(C.Structure):
    _fields_ = [("normal", Vec2), ("rollingImpulse", c_float), ("points", ManifoldPoint * 2), ("pointCount", c_int)]


class WorldDef(C.Structure):
    _fields_ = [("gravity", Vec2), ("restitutionThreshold", c_float), ("hitEventThreshold", c_float),
                ("contactHertz", c_float), ("contactDampingRatio", c_float), ("maxContactPushSpeed", c_float),
                ("maximumLinearSpeed", c_float), ("frictionCallback", c_void_p), ("restitutionCallback", c_void_p),
                ("enableSleep", c_bool), ("enableContinuous", c_bool), ("workerCount", c_int),
                ("enqueueTask", c_void_p), ("finishTask", c_void_p), ("userTaskContext", c_void_p),
                ("userData", c_void_p), ("internalValue", c_int)]


class BodyDef(C.Structure):
    _fields_ = [("type", c_int), ("position", Vec2), ("rotation", Rot), ("linearVelocity", Vec2),
                ("angularVelocity", c_float), ("linearDamping", c_float), ("angularDamping", c_float),
                ("gravityScale", c_float), ("sleepThreshold", c_float), ("name", C.c_char_p), ("userData", c_void_p),
                ("enableSleep", c_bool), ("isAwake", c_bool), ("fixedRotation", c_bool), ("isBullet", c_bool),
                ("isEnabled", c_bool), ("allowFastRotation", c_bool), ("internalValue", c_int)]


class Filter(C.Structure):
    _fields_ = [("categoryBits", C.c_uint64), ("maskBits", C.c_uint64), ("groupIndex", c_int)]


class SurfaceMaterial(C.Structure):
    _fields_ = [("friction", c_float), ("restitution", c_float), ("rollingResistance", c_float),
                ("tangentSpeed", c_float), ("userMaterialId", c_int), ("customColor", C.c_uint32)]


class ShapeDef(C.Structure):
    _fields_ = [("userData", c_void_p), ("material", SurfaceMaterial), ("density", c_float), ("filter", Filter),
                ("isSensor", c_bool), ("enableSensorEvents", c_bool), ("enableContactEvents", c_bool),
                ("enableHitEvents", c_bool), ("enablePreSolveEvents", c_bool), ("invokeContactCreation", c_bool),
                ("updateBodyMass", c_bool), ("internalValue", c_int)]


class RevoluteJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("localAnchorA", Vec2), ("localAnchorB", Vec2),
                ("referenceAngle", c_float), ("targetAngle", c_float), ("enableSpring", c_bool), ("hertz", c_float),
                ("dampingRatio", c_float), ("enableLimit", c_bool), ("lowerAngle", c_float), ("upperAngle", c_float),
                ("enableMotor", c_bool), ("maxMotorTorque", c_float), ("motorSpeed", c_float), ("drawSize", c_float),
                ("collideConnected", c_bool), ("userData", c_void_p), ("internalValue", c_int)]


class MassData(C.Structure):
    _fields_ = [("mass", c_float), ("center", Vec2), ("rotationalInertia", c_float)]


class ChainSegment(C.Structure):
    _fields_ = [("ghost1", Vec2), ("segment", Segment), ("ghost2", Vec2), ("chainId", c_int)]


class DistanceJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("localAnchorA", Vec2), ("localAnchorB", Vec2),
                ("length", c_float), ("enableSpring", c_bool), ("hertz", c_float), ("dampingRatio", c_float),
                ("enableLimit", c_bool), ("minLength", c_float), ("maxLength", c_float), ("enableMotor", c_bool),
                ("maxMotorForce", c_float), ("motorSpeed", c_float), ("collideConnected", c_bool),
                ("userData", c_void_p), ("internalValue", c_int)]


class MotorJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("linearOffset", Vec2), ("angularOffset", c_float),
                ("maxForce", c_float), ("maxTorque", c_float), ("correctionFactor", c_float),
                ("collideConnected", c_bool), ("userData", c_void_p), ("internalValue", c_int)]


class MouseJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("target", Vec2), ("hertz", c_float),
                ("dampingRatio", c_float), ("maxForce", c_float), ("collideConnected", c_bool),
                ("userData", c_void_p), ("internalValue", c_int)]


class FilterJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("userData", c_void_p), ("internalValue", c_int)]


class PrismaticJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("localAnchorA", Vec2), ("localAnchorB", Vec2),
                ("localAxisA", Vec2), ("referenceAngle", c_float), ("targetTranslation", c_float),
                ("enableSpring", c_bool), ("hertz", c_float), ("dampingRatio", c_float), ("enableLimit", c_bool),
                ("lowerTranslation", c_float), ("upperTranslation", c_float), ("enableMotor", c_bool),
                ("maxMotorForce", c_float), ("motorSpeed", c_float), ("collideConnected", c_bool),
                ("userData", c_void_p), ("internalValue", c_int)]


class WeldJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("localAnchorA", Vec2), ("localAnchorB", Vec2),
                ("referenceAngle", c_float), ("linearHertz", c_float), ("angularHertz", c_float),
                ("linearDampingRatio", c_float), ("angularDampingRatio", c_float), ("collideConnected", c_bool),
                ("userData", c_void_p), ("internalValue", c_int)]


class WheelJointDef(C.Structure):
    _fields_ = [("bodyIdA", BodyId), ("bodyIdB", BodyId), ("localAnchorA", Vec2), ("localAnchorB", Vec2),
                ("localAxisA", Vec2), ("enableSpring", c_bool), ("hertz", c_float), ("dampingRatio", c_float),
                ("enableLimit", c_bool), ("lowerTranslation", c_float), ("upperTranslation", c_float),
                ("enableMotor", c_bool), ("maxMotorTorque", c_float), ("motorSpeed", c_float),
                ("collideConnected", c_bool), ("userData", c_void_p), ("internalValue", c_int)]


class ChainId(C.Structure):
    _fields_ = [("index1", C.c_int32), ("world0", C.c_uint16), ("generation", C.c_uint16)]


class ChainDef(C.Structure):
    _fields_ = [("userData", c_void_p), ("points", C.POINTER(Vec2)), ("count", c_int),
                ("materials", C.POINTER(SurfaceMaterial)), ("materialCount", c_int), ("filter", Filter),
                ("isLoop", c_bool), ("enableSensorEvents", c_bool), ("internalValue", c_int)]


class QueryFilter(C.Structure):
    _fields_ = [("categoryBits", C.c_uint64), ("maskBits", C.c_uint64)]


class RayResult(C.Structure):
    _fields_ = [("shapeId", ShapeId), ("point", Vec2), ("normal", Vec2), ("fraction", c_float), ("nodeVisits", c_int),
                ("leafVisits", c_int), ("hit", c_bool)]


class TreeStats(C.Structure):
    _fields_ = [("nodeVisits", c_int), ("leafVisits", c_int)]


OverlapResultFcn = C.CFUNCTYPE(c_bool, ShapeId, c_void_p)
CastResultFcn = C.CFUNCTYPE(c_float, ShapeId, Vec2, Vec2, c_float, c_void_p)


# b2DebugDraw (types.h:1383-1460); b2HexColor is an int-sized enum
DrawPolygonFcn = C.CFUNCTYPE(None, C.POINTER(Vec2), c_int, c_int, c_void_p)
DrawSolidPolygonFcn = C.CFUNCTYPE(None, Transform, C.POINTER(Vec2), c_int, c_float, c_int, c_void_p)
DrawCircleFcn = C.CFUNCTYPE(None, Vec2, c_float, c_int, c_void_p)
DrawSolidCircleFcn = C.CFUNCTYPE(None, Transform, c_float, c_int, c_void_p)
DrawSolidCapsuleFcn = C.CFUNCTYPE(None, Vec2, Vec2, c_float, c_int, c_void_p)
DrawSegmentFcn = C.CFUNCTYPE(None, Vec2, Vec2, c_int, c_void_p)
DrawTransformFcn = C.CFUNCTYPE(None, Transform, c_void_p)
DrawPointFcn = C.CFUNCTYPE(None, Vec2, c_float, c_int, c_void_p)
DrawStringFcn = C.CFUNCTYPE(None, Vec2, C.c_char_p, c_int, c_void_p)
DEBUG_DRAW_OPTIONS = ("useDrawingBounds", "drawShapes", "drawJoints", "drawJointExtras", "drawBounds", "drawMass", "drawBodyNames",
                      "drawContacts", "drawGraphColors", "drawContactNormals", "drawContactImpulses", "drawContactFeatures",
                      "drawFrictionImpulses", "drawIslands")


class DebugDraw(C.Structure):
    _fields_ = ([("DrawPolygonFcn", DrawPolygonFcn), ("DrawSolidPolygonFcn", DrawSolidPolygonFcn), ("DrawCircleFcn", DrawCircleFcn),
                 ("DrawSolidCircleFcn", DrawSolidCircleFcn), ("DrawSolidCapsuleFcn", DrawSolidCapsuleFcn),
                 ("DrawSegmentFcn", DrawSegmentFcn), ("DrawTransformFcn", DrawTransformFcn), ("DrawPointFcn", DrawPointFcn),
                 ("DrawStringFcn", DrawStringFcn), ("drawingBounds", AABB)] +
                [(name, c_bool) for name in DEBUG_DRAW_OPTIONS] + [("context", c_void_p)])


class ExplosionDef(C.Structure):
    _fields_ = [("maskBits", C.c_uint64), ("position", Vec2), ("radius", c_float), ("falloff", c_float),
                ("impulsePerLength", c_float)]


class Counters(C.Structure):
    _fields_ = [("bodyCount", c_int), ("shapeCount", c_int), ("contactCount", c_int), ("jointCount", c_int),
                ("islandCount", c_int), ("stackUsed", c_int), ("staticTreeHeight", c_int), ("treeHeight", c_int),
                ("byteCount", c_int), ("taskCount", c_int), ("colorCounts", c_int * 12)]


class Profile(C.Structure):  # types.h:466-490
    _fields_ = [(n, c_float) for n in ("step", "pairs", "collide", "solve", "mergeIslands", "prepareStages", "solveConstraints",
                                       "prepareConstraints", "integrateVelocities", "warmStart", "solveImpulses",
                                       "integratePositions", "relaxImpulses", "applyRestitution", "storeImpulses", "splitIslands",
                                       "transforms", "hitEvents", "refit", "bullets", "sleepIslands", "sensors")]


class BodyMoveEvent(C.Structure):
    _fields_ = [("transform", Transform), ("bodyId", BodyId), ("userData", c_void_p), ("fellAsleep", c_bool)]


class BodyEvents(C.Structure):
    _fields_ = [("moveEvents", C.POINTER(BodyMoveEvent)), ("moveCount", c_int)]


class ContactBeginTouchEvent(C.Structure):
    _fields_ = [("shapeIdA", ShapeId), ("shapeIdB", ShapeId), ("manifold", Manifold)]


class ContactEndTouchEvent(C.Structure):
    _fields_ = [("shapeIdA", ShapeId), ("shapeIdB", ShapeId)]


class ContactHitEvent(C.Structure):
    _fields_ = [("shapeIdA", ShapeId), ("shapeIdB", ShapeId), ("point", Vec2), ("normal", Vec2),
                ("approachSpeed", c_float)]


class ContactEvents(C.Structure):
    _fields_ = [("beginEvents", C.POINTER(ContactBeginTouchEvent)), ("endEvents", C.POINTER(ContactEndTouchEvent)),
                ("hitEvents", C.POINTER(ContactHitEvent)), ("beginCount", c_int), ("endCount", c_int),
                ("hitCount", c_int)]


class SensorEvent(C.Structure):
    _fields_ = [("sensorShapeId", ShapeId), ("visitorShapeId", ShapeId)]


class SensorEvents(C.Structure):
    _fields_ = [("beginEvents", C.POINTER(SensorEvent)), ("endEvents", C.POINTER(SensorEvent)),
                ("beginCount", c_int), ("endCount", c_int)]


# ---- introspection records (include/forge2d_b200_debug.h) ---------------------------------------------------------
class BodyRecord(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("id", "setIndex", "localIndex", "islandId", "islandPrev", "islandNext",
                                         "type", "headContactKey", "contactCount", "headShapeId", "flags")] + \
               [(n, c_float) for n in ("px", "py", "qc", "qs", "cx", "cy", "c0x", "c0y", "q0c", "q0s", "vx", "vy",
                                       "w", "sleepTime", "invMass", "invInertia", "minExtent", "maxExtent", "lcx",
                                       "lcy")]


class ContactRecord(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("id", "shapeIdA", "shapeIdB", "setIndex", "colorIndex", "localIndex",
                                         "flags", "simFlags", "pointCount", "id0", "id1", "islandId", "islandPrev",
                                         "islandNext", "prevKeyA", "nextKeyA", "prevKeyB", "nextKeyB",
                                         "bodySimIndexA", "bodySimIndexB")] + \
               [("nx", c_float), ("ny", c_float)] + \
               [(n, c_float * 2) for n in ("sep", "ni", "ti", "tni", "nv", "ax", "ay", "bx", "by", "px", "py")] + \
               [("friction", c_float), ("restitution", c_float), ("rollingImpulse", c_float)]


class IslandRecord(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("id", "setIndex", "localIndex", "headBody", "tailBody", "bodyCount",
                                         "headContact", "tailContact", "contactCount", "headJoint", "tailJoint",
                                         "jointCount", "parentIsland", "constraintRemoveCount")]


class ShapeRecord(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("id", "bodyId", "proxyKey", "type", "enlarged")] + \
               [("aabb", c_float * 4), ("fat", c_float * 4)]


class TreeLeafRecord(C.Structure):
    _fields_ = [("proxyId", C.c_int32), ("depth", C.c_int32), ("enlargedAncestors", C.c_int32), ("box", c_float * 4)]


class JointRecord(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("id", "type", "setIndex", "colorIndex", "localIndex", "bodyIdA", "bodyIdB",
                                         "islandId")] + [("impulse", c_float * 6)]


# name -> (restype, argtypes): the b2* subset declared in include/forge2d_b200.h
B2_FUNCTIONS = {
    "b2DefaultWorldDef": (WorldDef, []),
    "b2DefaultBodyDef": (BodyDef, []),
    "b2DefaultFilter": (Filter, []),
    "b2DefaultShapeDef": (ShapeDef, []),
    "b2DefaultSurfaceMaterial": (SurfaceMaterial, []),
    "b2DefaultRevoluteJointDef": (RevoluteJointDef, []),
    "b2MakeBox": (Polygon, [c_float, c_float]),
    "b2MakeSquare": (Polygon, [c_float]),
    "b2MakeOffsetRoundedBox": (Polygon, [c_float, c_float, Vec2, Rot, c_float]),
    "b2MakePolygon": (Polygon, [C.POINTER(Hull), c_float]),
    "b2ComputeHull": (Hull, [C.POINTER(Vec2), c_int]),
    "b2CreateWorld": (WorldId, [C.POINTER(WorldDef)]),
    "b2DestroyWorld": (None, [WorldId]),
    "b2World_IsValid": (c_bool, [WorldId]),
    "b2World_Step": (None, [WorldId, c_float, c_int]),
    "b2World_GetBodyEvents": (BodyEvents, [WorldId]),
    "b2World_GetSensorEvents": (SensorEvents, [WorldId]),
    "b2World_GetContactEvents": (ContactEvents, [WorldId]),
    "b2World_EnableSleeping": (None, [WorldId, c_bool]),
    "b2World_IsSleepingEnabled": (c_bool, [WorldId]),
    "b2World_EnableContinuous": (None, [WorldId, c_bool]),
    "b2World_IsContinuousEnabled": (c_bool, [WorldId]),
    "b2World_SetGravity": (None, [WorldId, Vec2]),
    "b2World_GetGravity": (Vec2, [WorldId]),
    "b2World_EnableWarmStarting": (None, [WorldId, c_bool]),
    "b2World_GetCounters": (Counters, [WorldId]),
    "b2World_GetProfile": (Profile, [WorldId]),
    "b2World_GetAwakeBodyCount": (c_int, [WorldId]),
    "b2CreateBody": (BodyId, [WorldId, C.POINTER(BodyDef)]),
    "b2Body_IsValid": (c_bool, [BodyId]),
    "b2Body_GetType": (c_int, [BodyId]),
    "b2Body_GetPosition": (Vec2, [BodyId]),
    "b2Body_GetRotation": (Rot, [BodyId]),
    "b2Body_GetTransform": (Transform, [BodyId]),
    "b2Body_GetLinearVelocity": (Vec2, [BodyId]),
    "b2Body_GetAngularVelocity": (c_float, [BodyId]),
    "b2Body_SetLinearVelocity": (None, [BodyId, Vec2]),
    "b2Body_SetAngularVelocity": (None, [BodyId, c_float]),
    "b2Body_GetMass": (c_float, [BodyId]),
    "b2Body_GetRotationalInertia": (c_float, [BodyId]),
    "b2Body_GetLocalCenterOfMass": (Vec2, [BodyId]),
    "b2Body_GetWorldCenterOfMass": (Vec2, [BodyId]),
    "b2Body_IsAwake": (c_bool, [BodyId]),
    "b2Body_GetShapeCount": (c_int, [BodyId]),
    "b2Body_GetContactCapacity": (c_int, [BodyId]),
    "b2CreateCircleShape": (ShapeId, [BodyId, C.POINTER(ShapeDef), C.POINTER(Circle)]),
    "b2CreateSegmentShape": (ShapeId, [BodyId, C.POINTER(ShapeDef), C.POINTER(Segment)]),
    "b2CreateCapsuleShape": (ShapeId, [BodyId, C.POINTER(ShapeDef), C.POINTER(Capsule)]),
    "b2CreatePolygonShape": (ShapeId, [BodyId, C.POINTER(ShapeDef), C.POINTER(Polygon)]),
    "b2Shape_IsValid": (c_bool, [ShapeId]),
    "b2Shape_GetBody": (BodyId, [ShapeId]),
    "b2Shape_GetAABB": (AABB, [ShapeId]),
    "b2CreateRevoluteJoint": (JointId, [WorldId, C.POINTER(RevoluteJointDef)]),
    "b2Joint_IsValid": (c_bool, [JointId]),
    # ---- the rest of the body / shape / joint surface (f2d_capi_ext.inl)
    "b2DestroyBody": (None, [BodyId]),
    "b2Body_GetLocalPoint": (Vec2, [BodyId, Vec2]),
    "b2Body_GetWorldPoint": (Vec2, [BodyId, Vec2]),
    "b2Body_GetLocalVector": (Vec2, [BodyId, Vec2]),
    "b2Body_GetWorldVector": (Vec2, [BodyId, Vec2]),
    "b2Body_SetTransform": (None, [BodyId, Vec2, Rot]),
    "b2Body_ApplyForce": (None, [BodyId, Vec2, Vec2, c_bool]),
    "b2Body_ApplyForceToCenter": (None, [BodyId, Vec2, c_bool]),
    "b2Body_ApplyTorque": (None, [BodyId, c_float, c_bool]),
    "b2Body_ApplyLinearImpulse": (None, [BodyId, Vec2, Vec2, c_bool]),
    "b2Body_ApplyLinearImpulseToCenter": (None, [BodyId, Vec2, c_bool]),
    "b2Body_ApplyAngularImpulse": (None, [BodyId, c_float, c_bool]),
    "b2Body_SetType": (None, [BodyId, c_int]),
    "b2Body_SetName": (None, [BodyId, C.c_char_p]),
    "b2Body_GetName": (C.c_char_p, [BodyId]),
    "b2Body_SetUserData": (None, [BodyId, c_void_p]),
    "b2Body_GetUserData": (c_void_p, [BodyId]),
    "b2Body_SetMassData": (None, [BodyId, MassData]),
    "b2Body_GetMassData": (MassData, [BodyId]),
    "b2Body_ApplyMassFromShapes": (None, [BodyId]),
    "b2Body_SetLinearDamping": (None, [BodyId, c_float]),
    "b2Body_GetLinearDamping": (c_float, [BodyId]),
    "b2Body_SetAngularDamping": (None, [BodyId, c_float]),
    "b2Body_GetAngularDamping": (c_float, [BodyId]),
    "b2Body_SetGravityScale": (None, [BodyId, c_float]),
    "b2Body_GetGravityScale": (c_float, [BodyId]),
    "b2Body_SetAwake": (None, [BodyId, c_bool]),
    "b2Body_IsEnabled": (c_bool, [BodyId]),
    "b2Body_IsSleepEnabled": (c_bool, [BodyId]),
    "b2Body_SetSleepThreshold": (None, [BodyId, c_float]),
    "b2Body_GetSleepThreshold": (c_float, [BodyId]),
    "b2Body_EnableSleep": (None, [BodyId, c_bool]),
    "b2Body_Disable": (None, [BodyId]),
    "b2Body_Enable": (None, [BodyId]),
    "b2Body_SetFixedRotation": (None, [BodyId, c_bool]),
    "b2Body_IsFixedRotation": (c_bool, [BodyId]),
    "b2Body_SetBullet": (None, [BodyId, c_bool]),
    "b2Body_IsBullet": (c_bool, [BodyId]),
    "b2Body_EnableContactEvents": (None, [BodyId, c_bool]),
    "b2Body_EnableHitEvents": (None, [BodyId, c_bool]),
    "b2Body_GetWorld": (WorldId, [BodyId]),
    "b2Body_GetShapes": (c_int, [BodyId, C.POINTER(ShapeId), c_int]),
    "b2Body_GetJointCount": (c_int, [BodyId]),
    "b2Body_GetJoints": (c_int, [BodyId, C.POINTER(JointId), c_int]),
    "b2DestroyShape": (None, [ShapeId, c_bool]),
    "b2Shape_GetWorld": (WorldId, [ShapeId]),
    "b2Shape_SetUserData": (None, [ShapeId, c_void_p]),
    "b2Shape_GetUserData": (c_void_p, [ShapeId]),
    "b2Shape_IsSensor": (c_bool, [ShapeId]),
    "b2Shape_TestPoint": (c_bool, [ShapeId, Vec2]),
    "b2Shape_SetDensity": (None, [ShapeId, c_float, c_bool]),
    "b2Shape_GetDensity": (c_float, [ShapeId]),
    "b2Shape_SetFriction": (None, [ShapeId, c_float]),
    "b2Shape_GetFriction": (c_float, [ShapeId]),
    "b2Shape_SetRestitution": (None, [ShapeId, c_float]),
    "b2Shape_GetRestitution": (c_float, [ShapeId]),
    "b2Shape_GetFilter": (Filter, [ShapeId]),
    "b2Shape_SetFilter": (None, [ShapeId, Filter]),
    "b2Shape_EnableSensorEvents": (None, [ShapeId, c_bool]),
    "b2Shape_AreSensorEventsEnabled": (c_bool, [ShapeId]),
    "b2Shape_EnableContactEvents": (None, [ShapeId, c_bool]),
    "b2Shape_AreContactEventsEnabled": (c_bool, [ShapeId]),
    "b2Shape_EnablePreSolveEvents": (None, [ShapeId, c_bool]),
    "b2Shape_ArePreSolveEventsEnabled": (c_bool, [ShapeId]),
    "b2Shape_EnableHitEvents": (None, [ShapeId, c_bool]),
    "b2Shape_AreHitEventsEnabled": (c_bool, [ShapeId]),
    "b2Shape_GetType": (c_int, [ShapeId]),
    "b2Shape_GetCircle": (Circle, [ShapeId]),
    "b2Shape_GetSegment": (Segment, [ShapeId]),
    "b2Shape_GetChainSegment": (ChainSegment, [ShapeId]),
    "b2Shape_GetCapsule": (Capsule, [ShapeId]),
    "b2Shape_GetPolygon": (Polygon, [ShapeId]),
    "b2DefaultDistanceJointDef": (DistanceJointDef, []),
    "b2DefaultMotorJointDef": (MotorJointDef, []),
    "b2DefaultMouseJointDef": (MouseJointDef, []),
    "b2DefaultFilterJointDef": (FilterJointDef, []),
    "b2DefaultPrismaticJointDef": (PrismaticJointDef, []),
    "b2DefaultWeldJointDef": (WeldJointDef, []),
    "b2DefaultWheelJointDef": (WheelJointDef, []),
    "b2DefaultExplosionDef": (ExplosionDef, []),
    "b2CreateDistanceJoint": (JointId, [WorldId, C.POINTER(DistanceJointDef)]),
    "b2CreateMotorJoint": (JointId, [WorldId, C.POINTER(MotorJointDef)]),
    "b2CreateMouseJoint": (JointId, [WorldId, C.POINTER(MouseJointDef)]),
    "b2CreateFilterJoint": (JointId, [WorldId, C.POINTER(FilterJointDef)]),
    "b2CreatePrismaticJoint": (JointId, [WorldId, C.POINTER(PrismaticJointDef)]),
    "b2CreateWeldJoint": (JointId, [WorldId, C.POINTER(WeldJointDef)]),
    "b2CreateWheelJoint": (JointId, [WorldId, C.POINTER(WheelJointDef)]),
    "b2DestroyJoint": (None, [JointId]),
    "b2Joint_GetType": (c_int, [JointId]),
    "b2Joint_GetBodyA": (BodyId, [JointId]),
    "b2Joint_GetBodyB": (BodyId, [JointId]),
    "b2Joint_GetWorld": (WorldId, [JointId]),
    "b2Joint_GetLocalAnchorA": (Vec2, [JointId]),
    "b2Joint_GetLocalAnchorB": (Vec2, [JointId]),
    "b2Joint_SetLocalAnchorA": (None, [JointId, Vec2]),
    "b2Joint_SetLocalAnchorB": (None, [JointId, Vec2]),
    "b2Joint_SetCollideConnected": (None, [JointId, c_bool]),
    "b2Joint_GetCollideConnected": (c_bool, [JointId]),
    "b2Joint_SetUserData": (None, [JointId, c_void_p]),
    "b2Joint_GetUserData": (c_void_p, [JointId]),
    "b2Joint_WakeBodies": (None, [JointId]),
    "b2Joint_GetConstraintForce": (Vec2, [JointId]),
    "b2Joint_GetConstraintTorque": (c_float, [JointId]),
    "b2DefaultChainDef": (ChainDef, []),
    "b2CreateChain": (ChainId, [BodyId, C.POINTER(ChainDef)]),
    "b2DestroyChain": (None, [ChainId]),
    "b2Chain_IsValid": (c_bool, [ChainId]),
    "b2Chain_GetWorld": (WorldId, [ChainId]),
    "b2Chain_GetSegmentCount": (c_int, [ChainId]),
    "b2Chain_GetSegments": (c_int, [ChainId, C.POINTER(ShapeId), c_int]),
    "b2Chain_SetFriction": (None, [ChainId, c_float]),
    "b2Chain_GetFriction": (c_float, [ChainId]),
    "b2Chain_SetRestitution": (None, [ChainId, c_float]),
    "b2Chain_GetRestitution": (c_float, [ChainId]),
    "b2DefaultQueryFilter": (QueryFilter, []),
    "b2World_OverlapAABB": (TreeStats, [WorldId, AABB, QueryFilter, OverlapResultFcn, c_void_p]),
    "b2World_CastRay": (TreeStats, [WorldId, Vec2, Vec2, QueryFilter, CastResultFcn, c_void_p]),
    "b2World_CastRayClosest": (RayResult, [WorldId, Vec2, Vec2, QueryFilter]),
    "b2World_Explode": (None, [WorldId, C.POINTER(ExplosionDef)]),
    "b2DefaultDebugDraw": (DebugDraw, []),
    "b2World_Draw": (None, [WorldId, C.POINTER(DebugDraw)]),
    "b2World_SetCustomFilterCallback": (None, [WorldId, c_void_p, c_void_p]),
    "b2World_SetPreSolveCallback": (None, [WorldId, c_void_p, c_void_p]),
}

# per-joint-type accessors (generated table; names resolved to the classes above)
from ._abi_joints import JOINT_ACCESSORS as _JA  # noqa: E402

_NAMES = {"JointId": JointId, "Vec2": Vec2, "c_float": c_float, "c_bool": c_bool, None: None}
for _name, (_res, _args) in _JA.items():
    B2_FUNCTIONS[_name] = (_NAMES[_res], [_NAMES[a] for a in _args])

# the additive f2d* extension (product + emulation only)
F2D_FUNCTIONS = {
    "f2dBatch_Create": (c_void_p, [WorldId, c_int]),
    "f2dBatch_Destroy": (None, [c_void_p]),
    "f2dBatch_Step": (None, [c_void_p, c_float, c_int]),
    "f2dBatch_StepN": (None, [c_void_p, c_float, c_int, c_int]),
    "f2dBatch_Synchronize": (None, [c_void_p]),
    "f2dBatch_SetLaunchConfig": (c_int, [c_void_p, c_int, c_int]),
    "f2dBatch_GetWorldCount": (c_int, [c_void_p]),
    "f2dBatch_GetBodyEvents": (c_int, [c_void_p, C.POINTER(BodyMoveEvent), c_int, C.POINTER(c_int)]),
    "f2dBatch_DownloadWorld": (None, [c_void_p, c_int, WorldId]),
    "f2dBatch_ReadBodyEvents": (c_int, [c_void_p, c_int, C.POINTER(C.POINTER(BodyMoveEvent)), C.POINTER(C.POINTER(c_int))]),
    "f2dBatch_StepAndReadBodyEvents": (c_int, [c_void_p, c_float, c_int, c_int, C.POINTER(C.POINTER(BodyMoveEvent)),
                                               C.POINTER(C.POINTER(c_int))]),
    "f2dBatch_SetGravity": (None, [c_void_p, c_void_p, c_int]),
    "f2dBatch_EventRecord": (None, [c_void_p, c_int]),
    "f2dBatch_EventElapsedMs": (c_float, [c_void_p, c_int, c_int]),
    "f2dBatch_GetWorldBytes": (C.c_ulonglong, [c_void_p]),
    "f2dSetDevice": (c_int, [c_int]),
    "f2dHostAlloc": (c_void_p, [C.c_ulonglong]),
    "f2dHostFree": (None, [c_void_p]),
    "f2dBatch_GetErrorFlags": (C.c_uint32, [c_void_p]),
    "f2dBatch_CreateFromWorlds": (c_void_p, [C.POINTER(WorldId), c_int]),
    "f2dBatch_GetWorldErrors": (c_int, [c_void_p, C.POINTER(C.c_uint32), c_int]),
    "f2dBatch_GetGrowthCount": (c_int, [c_void_p]),
    "f2dBatch_SetGangMode": (None, [c_void_p, c_int]),
    "f2dGetTransferBytes": (None, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "f2dBatch_StepPipelined": (c_int, [c_void_p, C.c_float, c_int, c_int, c_int, C.POINTER(c_void_p), C.POINTER(C.POINTER(c_int))]),
    "f2dBatch_FlushPipelined": (c_int, [c_void_p, C.POINTER(c_void_p), C.POINTER(C.POINTER(c_int))]),
    "f2dBatch_TranslateWorlds": (None, [c_void_p, c_void_p, c_int]),
    "f2dHasDevice": (c_int, []),
    "f2dGetLastError": (C.c_char_p, []),
    "f2dClearLastError": (None, []),
    "f2dWorld_SetLaunchMode": (None, [WorldId, c_int]),
    "f2dWorld_GetErrorFlags": (C.c_uint32, [WorldId]),
    "f2dWorld_GetKernelLaunchCount": (C.c_longlong, []),
    "f2dWorld_GetLastStepTimes": (None, [WorldId, C.POINTER(c_float)]),
    "f2dWorld_EnablePhaseTiming": (None, [WorldId, c_bool]),
    "f2dWorld_GetStepInfo": (c_int, [WorldId, C.POINTER(c_int), c_int]),
    "f2dWorld_EnableProfile": (None, [WorldId, c_bool]),
    "f2dWorld_EnablePairClassBinning": (None, [WorldId, c_bool]),
    "f2dWorld_ReadProfile": (c_int, [WorldId, C.POINTER(C.c_ulonglong), c_int]),
    "f2dWorld_StepAsync": (None, [WorldId, c_float, c_int]),
    "f2dWorld_Synchronize": (None, [WorldId]),
}

# introspection: product uses the f2dDebug_ prefix, the reference tap (oracle/tap.c) the tap_ prefix
DEBUG_FUNCTIONS = {
    "awake_order": ("f2dDebug_AwakeOrder", "tap_awake_order", c_int, [WorldId, C.POINTER(c_int), c_int]),
    "move_array": ("f2dDebug_MoveArray", "tap_move_array", c_int, [WorldId, C.POINTER(c_int), c_int]),
    "bodies": ("f2dDebug_Bodies", "tap_bodies", c_int, [WorldId, C.POINTER(BodyRecord), c_int]),
    "contacts": ("f2dDebug_Contacts", "tap_contacts", c_int, [WorldId, C.POINTER(ContactRecord), c_int]),
    "islands": ("f2dDebug_Islands", "tap_islands", c_int, [WorldId, C.POINTER(IslandRecord), c_int]),
    "shapes": ("f2dDebug_Shapes", "tap_shapes", c_int, [WorldId, C.POINTER(ShapeRecord), c_int]),
    "tree": ("f2dDebug_Tree", "tap_tree", c_int, [WorldId, c_int, C.POINTER(TreeLeafRecord), c_int]),
    "joints": ("f2dDebug_Joints", "tap_joints", c_int, [WorldId, C.POINTER(JointRecord), c_int]),
    "color_counts": ("f2dDebug_ColorCounts", "tap_color_counts", None, [WorldId, C.POINTER(c_int), C.POINTER(c_int)]),
    "color_contacts": ("f2dDebug_ColorContacts", "tap_color_contacts", c_int, [WorldId, c_int, C.POINTER(c_int), c_int]),
    "awake_contacts": ("f2dDebug_AwakeContacts", "tap_awake_contacts", c_int, [WorldId, C.POINTER(c_int), c_int]),
    "awake_islands": ("f2dDebug_AwakeIslands", "tap_awake_islands", c_int, [WorldId, C.POINTER(c_int), c_int]),
}


class Library:
    """A loaded shared library exporting the b2* ABI, with typed entry points as attributes."""

    def __init__(self, path, kind):
        self.path = path
        self.kind = kind  # "product" | "emu" | "reference"
        self.dll = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
        self.missing = []
        for name, (res, args) in B2_FUNCTIONS.items():
            self._bind(name, name, res, args)
        if kind != "reference":
            for name, (res, args) in F2D_FUNCTIONS.items():
                self._bind(name, name, res, args)
        for attr, (prod, tap, res, args) in DEBUG_FUNCTIONS.items():
            self._bind("debug_" + attr, tap if kind == "reference" else prod, res, args)

    def _bind(self, attr, symbol, res, args):
        try:
            fn = getattr(self.dll, symbol)
        except AttributeError:
            self.missing.append(symbol)
            return
        fn.restype = res
        fn.argtypes = args
        setattr(self, attr, fn)
