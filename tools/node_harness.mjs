// node_harness.mjs — post one frame of Jobs through ts/gpuWorkerShim.ts (the `Worker` replacement) on a box with Node >= 18 and a
// B200, and diff the Results against the ctypes path.  One command:
//
//     (cd addon && npm install && npx node-gyp rebuild) && RM_REFERENCE=/path/to/cpu-raymarcher npx tsx tools/node_harness.mjs
//
// What it does (main.ts:318-321,444-490 in 40 lines): installs globalThis.Worker = GpuWorker, creates NUM_WORKERS workers,
// partitions the rows, posts every Job, awaits one message per worker, assembles the frame buffers; then runs
// `python tools/dump_frame.py` for the same job and compares the planes byte for byte.
// (In the build container there is no Node.js: the same flow is executed against the addon under tests/napi_mock.cc instead,
// tests/test_gpu_addon.py.)
import { execFileSync } from 'node:child_process';
import fs from 'node:fs';
import os from 'node:os';
import path from 'node:path';

const here = path.dirname(new URL(import.meta.url).pathname);
await import(path.join(here, '..', 'ts', 'gpuWorkerShim.ts'));     // sets globalThis.Worker

const NUM_WORKERS = 4, width = 640, height = 360;
const base = { width, height, time: 0, camera: { pitch: 0.1, yaw: 0.5 }, algorithm: 'sphere-tracer', scenePresetIndex: 2,
               accelerationStructure: 'BVH', overshootFactor: 1.2, stepSize: 0.1 };
const workers = Array.from({ length: NUM_WORKERS }, () => new Worker(new URL('file:///unused'), { type: 'module' }));
const rowsPerWorker = Math.ceil(height / NUM_WORKERS);
const depth = new Uint8ClampedArray(width * height), normal = new Uint8ClampedArray(3 * width * height);
const sdfEval = new Uint16Array(width * height), iters = new Uint16Array(width * height);
await Promise.all(workers.map((w, i) => new Promise((resolve) => {
  const yStart = Math.min(i * rowsPerWorker, height), yEnd = Math.min((i + 1) * rowsPerWorker, height);
  if (yStart >= yEnd) return resolve(null);
  const h = (e) => {
    w.removeEventListener('message', h);
    const r = e.data, o = r.yStart * width;                          // main.ts:461-468
    depth.set(r.depth, o); normal.set(r.normal, 3 * o); sdfEval.set(r.sdfEval, o); iters.set(r.iters, o);
    resolve(r);
  };
  w.addEventListener('message', h);
  w.postMessage({ ...base, yStart, yEnd });
})));
const tmp = fs.mkdtempSync(path.join(os.tmpdir(), 'rm-'));
const dump = path.join(tmp, 'frame.bin');
execFileSync('python', [path.join(here, 'dump_frame.py'), '--dump', dump, '--job', JSON.stringify({ ...base, yStart: 0, yEnd: height })], { stdio: 'inherit' });
const want = fs.readFileSync(dump);
const got = Buffer.concat([Buffer.from(depth.buffer), Buffer.from(normal.buffer), Buffer.from(sdfEval.buffer), Buffer.from(iters.buffer)]);
if (Buffer.compare(want, got) !== 0) { console.error('MISMATCH between the addon frame and the ctypes frame'); process.exit(1); }
console.log(`OK: ${NUM_WORKERS} band Jobs through GpuWorker == ctypes frame (${got.length} bytes)`);
