"""The reference's own behavioural tests for the step path, transliterated from Dart onto the Python mirror of its API
(forge2d_b200/api.py). Sources: packages/forge2d/test/api/world_test.dart:8-74, events_test.dart:17-148 (contact
begin/end/hit, body move events, fellAsleep), packages/forge2d/test/ffi/smoke_test.dart:13-51.

Run twice: on CPU against the host emulation of the step templates (host logic of the API layer), and with `-m gpu`
against the product library, where every World.step is a CUDA kernel launch."""
import pytest

from forge2d_b200 import api
from forge2d_b200.api import (BodyDef, BodyType, Circle, Polygon, ShapeDef, Vector2, World, WorldDef)


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    lib = request.getfixturevalue(request.param)
    api.initializeForge2D(lib)
    yield lib
    api._backend = None


@pytest.fixture
def world(backend):
    w = World()
    yield w
    if w.isValid:
        w.destroy()


def ground(world):
    body = world.createBody(BodyDef(position=Vector2(0, -1)))
    shape = body.createShape(Polygon.box(50, 1), ShapeDef(enableContactEvents=True))
    return body, shape


# ---- world_test.dart
def test_world_starts_valid_and_becomes_invalid_on_destroy(backend):
    w = World()
    assert w.isValid
    w.destroy()
    assert not w.isValid


def test_world_gravity_default_override_and_change(backend):
    w = World()
    assert w.gravity == Vector2(0, -10)
    w.gravity = Vector2(0, -3.5)
    assert w.gravity == Vector2(0, -3.5)
    w.destroy()
    w = World(gravity=Vector2(3, -1.5))
    assert w.gravity == Vector2(3, -1.5)
    w.destroy()


def test_definition_toggles_are_applied(backend):
    w = World(definition=WorldDef(enableSleep=False, enableContinuous=False))
    assert not w.sleepingEnabled and not w.continuousEnabled
    w.sleepingEnabled = True
    w.continuousEnabled = True
    assert w.sleepingEnabled and w.continuousEnabled
    w.destroy()


def test_destroying_a_world_destroys_its_bodies(backend):
    w = World()
    body = w.createBody(BodyDef(userData="payload"))
    assert body.isValid and body.userData == "payload"
    w.destroy()
    assert not body.isValid


def test_a_falling_box_lands_on_a_static_ground_box(world):
    world.createBody(BodyDef(position=Vector2(0, -1))).createShape(Polygon.box(50, 1))
    box = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 10)))
    box.createShape(Polygon.square(0.5))
    for _ in range(200):
        world.step(1 / 60)
    assert box.position.x == pytest.approx(0, abs=0.01)
    assert box.position.y == pytest.approx(0.5, abs=0.01)
    assert not box.isAwake


# ---- events_test.dart
def test_begin_and_end_events_fire_for_a_bouncing_contact(world):
    _, ground_shape = ground(world)
    ball = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 2)))
    ball_shape = ball.createShape(Circle(radius=0.5), ShapeDef(enableContactEvents=True, restitution=0.8))
    began, ended = [], []
    for _ in range(120):
        world.step(1 / 60)
        events = world.contactEvents
        began += events.begin
        ended += events.end
    assert began
    touching = {began[0].shapeA, began[0].shapeB}
    assert ball_shape in touching and ground_shape in touching
    assert began[0].points
    assert abs(began[0].normal.y) == pytest.approx(1, abs=0.01)
    assert ended  # the bouncy ball leaves the ground again


def test_hit_events_report_the_approach_speed(world):
    ground(world)
    world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 5))).createShape(
        Circle(radius=0.5), ShapeDef(enableContactEvents=True, enableHitEvents=True))
    hits = []
    for _ in range(120):
        world.step(1 / 60)
        hits += world.contactEvents.hit
    assert hits
    assert hits[0].approachSpeed > 1  # dropped from 5 m: above the default 1 m/s threshold


def test_no_events_without_enable_contact_events(world):
    world.createBody(BodyDef(position=Vector2(0, -1))).createShape(Polygon.box(50, 1), ShapeDef(enableContactEvents=False))
    world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 2))).createShape(
        Circle(radius=0.5), ShapeDef(enableContactEvents=False))
    for _ in range(120):
        world.step(1 / 60)
        assert world.contactEvents.begin == []


def test_only_moving_bodies_are_reported(world):
    ground(world)
    faller = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(10, 5)))
    faller.createShape(Circle(radius=0.5))
    world.step(1 / 60)
    events = world.bodyMoveEvents
    mine = [e for e in events if e.body == faller]
    assert len(mine) == 1
    assert mine[0].transform.p.y < 5
    assert mine[0].fellAsleep is False


def test_fell_asleep_is_reported_when_a_body_comes_to_rest(world):
    ground(world)
    world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 0.6))).createShape(Polygon.square(0.5))
    reported = False
    for _ in range(300):
        world.step(1 / 60)
        if any(e.fellAsleep for e in world.bodyMoveEvents):
            reported = True
            break
    assert reported


# ---- locked_world_test.dart: the locked flag only spans step()
def test_world_is_unlocked_after_step_and_creation_works(world):
    world.step(1 / 60)
    assert world.locked is False
    body = world.createBody(BodyDef(type=BodyType.dynamic))
    assert body.isValid


# ---- smoke_test.dart: stepping an empty world
def test_stepping_an_empty_world(world):
    for _ in range(5):
        world.step(1 / 60)
    assert world.bodyMoveEvents == []


# ---- events_test.dart:261-305 "simulation callbacks"
def test_the_custom_filter_can_disable_a_collision(world):
    ground(world)
    ball = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 3)))
    ball.createShape(Circle(radius=0.5), ShapeDef(userData="ghost"))
    world.customFilterCallback = lambda a, b: a.userData != "ghost" and b.userData != "ghost"
    for _ in range(120):
        world.step(1 / 60)
    assert ball.position.y < -1     # the ball ignored the ground and kept falling
    world.customFilterCallback = None


def test_the_pre_solve_callback_can_disable_a_contact(world):
    ground(world)
    ball = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 2)))
    ball.createShape(Circle(radius=0.5), ShapeDef(enablePreSolveEvents=True))
    called = []

    def pre_solve(shape_a, shape_b, normal):
        called.append((shape_a.key, shape_b.key, normal))
        return False

    world.preSolveCallback = pre_solve
    for _ in range(120):
        world.step(1 / 60)
    assert called
    assert ball.position.y < -1     # with every contact disabled the ball falls through the ground
    world.preSolveCallback = None


def test_removing_the_callbacks_restores_the_collision(world):
    ground(world)
    ball = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(0, 3)))
    ball.createShape(Circle(radius=0.5), ShapeDef(enablePreSolveEvents=True))
    world.customFilterCallback = lambda a, b: True
    world.preSolveCallback = lambda a, b, normal: True
    for _ in range(30):
        world.step(1 / 60)
    world.customFilterCallback = None
    world.preSolveCallback = None
    for _ in range(150):
        world.step(1 / 60)
    assert abs(ball.position.y - 0.5) < 0.05    # resting on the ground box (top at y = 0)


# ---- debug_draw_test.dart
class RecordingDebugDraw(api.DebugDraw):
    def __init__(self):
        self.solidPolygons, self.solidCircles, self.solidCapsules, self.segments, self.points, self.strings = [], [], [], [], [], []

    def drawSolidPolygon(self, transform, vertices, radius, color):
        self.solidPolygons.append(vertices)

    def drawSolidCircle(self, transform, radius, color):
        self.solidCircles.append(transform)

    def drawSolidCapsule(self, p1, p2, radius, color):
        self.solidCapsules.append((p1, p2, radius))

    def drawSegment(self, p1, p2, color):
        self.segments.append((p1, p2))

    def drawPoint(self, p, size, color):
        self.points.append(p)

    def drawString(self, p, text, color):
        self.strings.append(text)


def test_shapes_are_drawn_with_the_matching_primitives(world):
    body = world.createBody(BodyDef(position=Vector2(1, 2)))
    body.createShape(Polygon.square(0.5))
    body.createShape(Circle(radius=0.5))
    body.createShape(api.Capsule(center1=Vector2(0, 0), center2=Vector2(0, 1), radius=0.25))
    draw = RecordingDebugDraw()
    world.draw(draw)
    assert len(draw.solidPolygons) == 1 and len(draw.solidPolygons[0]) == 4
    assert len(draw.solidCircles) == 1 and abs(draw.solidCircles[0].p.x - 1) < 1e-5
    assert len(draw.solidCapsules) == 1 and abs(draw.solidCapsules[0][2] - 0.25) < 1e-6


def test_joints_are_drawn_when_enabled(world):
    anchor = world.createBody(BodyDef(position=Vector2(0, 5)))
    swinging = world.createBody(BodyDef(type=BodyType.dynamic, position=Vector2(2, 5)))
    swinging.createShape(Polygon.square(0.25))
    world.createRevoluteJoint(api.RevoluteJointDef(bodyA=anchor, bodyB=swinging))
    draw = RecordingDebugDraw()
    world.draw(draw)
    assert draw.segments
    draw.segments.clear()
    draw.drawJoints = False
    world.draw(draw)
    assert draw.segments == []


def test_body_names_are_drawn_when_enabled(world):
    world.createBody(BodyDef(name="labeled")).createShape(Circle(radius=1))
    draw = RecordingDebugDraw()
    draw.drawBodyNames = True
    world.draw(draw)
    assert "labeled" in draw.strings


def test_drawing_bounds_cull_far_away_shapes(world):
    world.createBody(BodyDef(position=Vector2(100, 100))).createShape(Polygon.square(0.5))
    draw = RecordingDebugDraw()
    draw.drawingBounds = ((-10, -10), (10, 10))
    world.draw(draw)
    assert draw.solidPolygons == []
