// gpuWorkerShim.ts — drop-in `Worker` replacement so src/main.ts runs unchanged on the B200 addon.
//
// main.ts (reference) does, per worker object (main.ts:319-321, 452-486):
//     const w = new Worker(new URL('./workers/raymarchWorker.ts', import.meta.url), {type: 'module'});
//     w.addEventListener('message', handler);   // one-shot per frame
//     w.postMessage(job);                       // Job of raymarchWorker.ts:10-22
// Install this class as globalThis.Worker (Electron preload / node + jsdom harness) before main.ts loads.
// The reference keeps scene / camera / shader selection (SceneManager, Camera are imported unchanged);
// only the march moves to the GPU.  The addon owns every GPU of the box (rm_pool_*): all NUM_WORKERS GpuWorker objects talk to
// that one pool, and the row-band Jobs of a frame share ONE render (the pool's frame cache).  Not executed in the build container
// (no Node.js there): tests/addon_harness.cc drives the same addon calls under a mock N-API runtime; tools/node_harness.mjs is
// the one-command check under real Node.
import { mat3, mat4, vec3 } from 'gl-matrix';
import { SceneManager } from '../../reference/src/util/sceneManager';
import { Camera } from '../../reference/src/util/camera';
import { Sphere } from '../../reference/src/util/primitives/sphere';
import { Box } from '../../reference/src/util/primitives/box';
import { Torus } from '../../reference/src/util/primitives/torus';
import { Mandelbulb } from '../../reference/src/util/primitives/mandelbulb';
import { Primitive } from '../../reference/src/util/primitives/primitive';
import { Round } from '../../reference/src/util/primitive_operations/round';
import { Twist } from '../../reference/src/util/primitive_operations/twist';
import { SmoothUnion } from '../../reference/src/util/primitive_operations/smoothUnion';
import { SmoothSubtraction } from '../../reference/src/util/primitive_operations/smoothSubstraction';
import { Repetition } from '../../reference/src/util/primitive_operations/repetition';
import { AnimatedTranslate } from '../../reference/src/util/primitive_operations/animatedTranslate';

// eslint-disable-next-line @typescript-eslint/no-var-requires
const addon = require('../addon/build/Release/rm_napi.node');

type Job = {
  width: number; height: number; time: number; yStart: number; yEnd: number;
  camera: { pitch: number; yaw: number }; algorithm: string; scenePresetIndex: number;
  accelerationStructure: string; overshootFactor?: number; stepSize?: number;
};

let uploadedKey = '';

const NODE_BYTES = 128; // sizeof(rm_op_node): i32 kind, i32 child[2], i32 prim, f64 p[4], f32 dir[4], f32 transform[16]

// Scene.objectSDFs -> rm_scene arrays: leaves into the primitive SoA, operators into rm_op_node records (pre-order).
function flattenObjects(index: number, objects: Primitive[]) {
  const type: number[] = [], w2l: number[] = [], params: number[] = [];
  const nodes: { kind: number; child: number[]; prim: number; p: number[]; dir: number[]; transform: ArrayLike<number> }[] = [];
  const visit = (o: Primitive): number => {
    const me = nodes.length;
    const rec = { kind: 0, child: [-1, -1], prim: -1, p: [0, 0, 0, 0], dir: [0, 0, 0, 0], transform: o.transform as Float32Array };
    nodes.push(rec);
    if (o instanceof Sphere || o instanceof Box || o instanceof Torus || o instanceof Mandelbulb) {
      rec.prim = type.length;
      w2l.push(...(o.transform as Float32Array));                    // Primitive.transform (primitive.ts:4,10)
      if (o instanceof Sphere) { type.push(0); params.push(o.radius, 0, 0, 0); }
      else if (o instanceof Box) { type.push(1); params.push(o.halfSize[0], o.halfSize[1], o.halfSize[2], 0); }
      else if (o instanceof Torus) { type.push(2); params.push(o.majorRadius, o.minorRadius, 0, 0); }
      else { type.push(3); params.push(o.power, o.iterations, o.enableAnimation ? 1 : 0, o.animationSpeed); }
    } else if (o instanceof Round) { rec.kind = 1; rec.p[0] = o.radius; rec.child[0] = visit(o.primitive); }
    else if (o instanceof Twist) { rec.kind = 2; rec.p[0] = o.twistAmount; rec.child[0] = visit(o.primitive); }
    else if (o instanceof SmoothUnion) { rec.kind = 3; rec.p[0] = o.smoothness; rec.child = [visit(o.prim1), visit(o.prim2)]; }
    else if (o instanceof SmoothSubtraction) { rec.kind = 4; rec.p[0] = o.smoothness; rec.child = [visit(o.prim1), visit(o.prim2)]; }
    else if (o instanceof Repetition) { rec.kind = 5; rec.p = [o.spacing[0], o.spacing[1], o.spacing[2], 0]; rec.child[0] = visit(o.primitive); }
    else if (o instanceof AnimatedTranslate) {
      rec.kind = 6; rec.p = [o.amplitude, o.speed, 0, 0]; rec.dir = [o.direction[0], o.direction[1], o.direction[2], 0];
      rec.child[0] = visit(o.primitive);
    } else throw new Error(`preset ${index}: ${o.constructor.name} is outside the B200 hot path (no CPU fallback)`);
    return me;
  };
  const roots = objects.map(visit);
  const flat = objects.every(o => o instanceof Sphere || o instanceof Box || o instanceof Torus || o instanceof Mandelbulb);
  const buf = new ArrayBuffer(NODE_BYTES * nodes.length), dv = new DataView(buf);
  nodes.forEach((n, i) => {
    const o = NODE_BYTES * i;
    dv.setInt32(o, n.kind, true); dv.setInt32(o + 4, n.child[0], true); dv.setInt32(o + 8, n.child[1], true); dv.setInt32(o + 12, n.prim, true);
    n.p.forEach((v, k) => dv.setFloat64(o + 16 + 8 * k, v, true));
    n.dir.forEach((v, k) => dv.setFloat32(o + 48 + 4 * k, v, true));
    for (let k = 0; k < 16; k++) dv.setFloat32(o + 64 + 4 * k, n.transform[k], true);
  });
  return {
    type: Uint8Array.from(type), worldToLocal: Float32Array.from(w2l), params: Float64Array.from(params),
    opNodes: flat ? undefined : new Uint8Array(buf), objectRoot: flat ? undefined : Int32Array.from(roots),
  };
}

function uploadPreset(index: number, accel: string) {
  const key = `${index}/${accel}`;
  if (key === uploadedKey) return;
  const scene = flattenObjects(index, SceneManager.getPreset(index).objects);
  addon.uploadScene({ ...scene, accel });   // rm_upload_scene compiles the operator trees and builds the BVH / octree natively
  uploadedKey = key;
}

export class GpuWorker {
  private listeners = new Set<(e: { data: unknown }) => void>();
  addEventListener(_t: 'message', h: (e: { data: unknown }) => void) { this.listeners.add(h); }
  removeEventListener(_t: 'message', h: (e: { data: unknown }) => void) { this.listeners.delete(h); }
  postMessage(job: Job) {
    uploadPreset(job.scenePresetIndex, job.accelerationStructure);
    const cam = new Camera();
    cam.setAngles(job.camera.pitch, job.camera.yaw);               // raymarchWorker.ts:39
    const rot3 = mat3.create();
    mat3.fromMat4(rot3, cam.getRotationMatrix(mat4.create()));      // raymarcher.ts:62-64
    const origin = vec3.create();
    cam.getPosition(origin);                                         // raymarcher.ts:66-67
    addon.render(job, { rot3, origin }).then((result: unknown) => {
      for (const h of Array.from(this.listeners)) h({ data: result }); // exactly one reply per request
    });
  }
  terminate() { /* nothing to do */ }
}

(globalThis as unknown as { Worker: unknown }).Worker = GpuWorker;
