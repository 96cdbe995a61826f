// forge2d_b200 — Dart side of the additive batch extension (include/forge2d_b200.h, `f2dBatch_*`).
//
// NOT compiled or run in this repository's CI: the build image has no Dart toolchain (DESIGN.md §1). It is the file a
// forge2d maintainer would add next to packages/forge2d/lib/src/ffi/box2d.g.dart: the externals use the same
// `@Native` mechanism and the same asset id as the generated Box2D bindings (box2d.g.dart:10), so they resolve in the
// library the build hook registers (INTEGRATION.md §1) — libforge2d_b200.so exports them next to the b2* symbols.
@ffi.DefaultAsset('package:forge2d/src/ffi/box2d.g.dart')
library;

import 'dart:ffi' as ffi;
import 'dart:typed_data';

import 'package:forge2d/src/api/world.dart';
import 'package:forge2d/src/ffi/box2d.g.dart' show b2BodyMoveEvent, b2Vec2, b2WorldId;

final class f2dBatch extends ffi.Opaque {}

@ffi.Native<ffi.Pointer<f2dBatch> Function(b2WorldId, ffi.Int)>()
external ffi.Pointer<f2dBatch> f2dBatch_Create(b2WorldId templateWorld, int count);

@ffi.Native<ffi.Void Function(ffi.Pointer<f2dBatch>)>()
external void f2dBatch_Destroy(ffi.Pointer<f2dBatch> batch);

@ffi.Native<ffi.Void Function(ffi.Pointer<f2dBatch>, ffi.Float, ffi.Int)>()
external void f2dBatch_Step(ffi.Pointer<f2dBatch> batch, double timeStep, int subStepCount);

@ffi.Native<
  ffi.Int Function(
    ffi.Pointer<f2dBatch>,
    ffi.Float,
    ffi.Int,
    ffi.Int,
    ffi.Pointer<ffi.Pointer<b2BodyMoveEvent>>,
    ffi.Pointer<ffi.Pointer<ffi.Int>>,
  )
>()
external int f2dBatch_StepAndReadBodyEvents(
  ffi.Pointer<f2dBatch> batch,
  double timeStep,
  int subStepCount,
  int maxBodiesPerWorld,
  ffi.Pointer<ffi.Pointer<b2BodyMoveEvent>> outEvents,
  ffi.Pointer<ffi.Pointer<ffi.Int>> outCounts,
);

@ffi.Native<ffi.Void Function(ffi.Pointer<f2dBatch>, ffi.Pointer<b2Vec2>, ffi.Int)>()
external void f2dBatch_SetGravity(ffi.Pointer<f2dBatch> batch, ffi.Pointer<b2Vec2> gravity, int count);

@ffi.Native<ffi.Void Function(ffi.Pointer<f2dBatch>, ffi.Int, b2WorldId)>()
external void f2dBatch_DownloadWorld(ffi.Pointer<f2dBatch> batch, int index, b2WorldId into);

@ffi.Native<ffi.Uint32 Function(ffi.Pointer<f2dBatch>)>()
external int f2dBatch_GetErrorFlags(ffi.Pointer<f2dBatch> batch);

/// `count` device-resident replicas of [template] stepped together, one thread block per world, sharded by world
/// across GPUs by running one isolate / process per GPU (no communication on the step path).
class WorldBatch {
  WorldBatch(World template, this.count) : _batch = f2dBatch_Create(template.id, count) {
    if (_batch == ffi.nullptr) {
      throw StateError('f2dBatch_Create failed (no CUDA device, or the template world has step callbacks)');
    }
  }

  final int count;
  final ffi.Pointer<f2dBatch> _batch;

  void step(double timeStep, {int subStepCount = 4}) => f2dBatch_Step(_batch, timeStep, subStepCount);

  /// One step of every world plus that step's body transforms. The returned view aliases pinned memory owned by the
  /// batch and stays valid until the next call (the ownership rule of `b2World_GetBodyEvents`).
  ({ffi.Pointer<b2BodyMoveEvent> events, Int32List counts}) stepAndReadBodyEvents(
    double timeStep,
    int maxBodiesPerWorld, {
    int subStepCount = 4,
    required ffi.Pointer<ffi.Pointer<b2BodyMoveEvent>> eventsOut,
    required ffi.Pointer<ffi.Pointer<ffi.Int>> countsOut,
  }) {
    f2dBatch_StepAndReadBodyEvents(_batch, timeStep, subStepCount, maxBodiesPerWorld, eventsOut, countsOut);
    return (events: eventsOut.value, counts: countsOut.value.cast<ffi.Int32>().asTypedList(count));
  }

  int get errorFlags => f2dBatch_GetErrorFlags(_batch);

  void destroy() => f2dBatch_Destroy(_batch);
}
