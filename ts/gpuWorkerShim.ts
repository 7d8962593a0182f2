// gpuWorkerShim.ts — drop-in `Worker` replacement so src/main.ts runs unchanged on the B200 addon.
//
// main.ts (reference) does, per worker object (main.ts:319-321, 452-486):
//     const w = new Worker(new URL('./workers/raymarchWorker.ts', import.meta.url), {type: 'module'});
//     w.addEventListener('message', handler);   // one-shot per frame
//     w.postMessage(job);                       // Job of raymarchWorker.ts:10-22
// Install this class as globalThis.Worker (Electron preload / node + jsdom harness) before main.ts loads.
// The reference keeps scene / camera / shader selection (SceneManager, Camera are imported unchanged);
// only the march moves to the GPU.  NOT executed in the build container (no Node.js there).
import { mat3, mat4, vec3 } from 'gl-matrix';
import { SceneManager } from '../../reference/src/util/sceneManager';
import { Camera } from '../../reference/src/util/camera';
import { Sphere } from '../../reference/src/util/primitives/sphere';
import { Box } from '../../reference/src/util/primitives/box';
import { Torus } from '../../reference/src/util/primitives/torus';

// eslint-disable-next-line @typescript-eslint/no-var-requires
const addon = require('../addon/build/Release/rm_napi.node');

type Job = {
  width: number; height: number; time: number; yStart: number; yEnd: number;
  camera: { pitch: number; yaw: number }; algorithm: string; scenePresetIndex: number;
  accelerationStructure: string; overshootFactor?: number; stepSize?: number;
};

let uploadedKey = '';

function uploadPreset(index: number, accel: string) {
  const key = `${index}/${accel}`;
  if (key === uploadedKey) return;
  const prims = SceneManager.getPreset(index).objects;
  const n = prims.length;
  const type = new Uint8Array(n), worldToLocal = new Float32Array(16 * n), params = new Float64Array(4 * n);
  prims.forEach((p, i) => {
    worldToLocal.set(p.transform as Float32Array, 16 * i);          // Primitive.transform (primitive.ts:4,10)
    if (p instanceof Sphere) { type[i] = 0; params[4 * i] = p.radius; }
    else if (p instanceof Box) { type[i] = 1; params.set([p.halfSize[0], p.halfSize[1], p.halfSize[2]], 4 * i); }
    else if (p instanceof Torus) { type[i] = 2; params[4 * i] = p.majorRadius; params[4 * i + 1] = p.minorRadius; }
    else throw new Error(`preset ${index}: ${p.constructor.name} is outside the B200 hot path (sphere/box/torus only)`);
  });
  addon.uploadScene({ type, worldToLocal, params, accel });        // rm_upload_scene builds the BVH / octree natively
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
