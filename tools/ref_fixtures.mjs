// ref_fixtures.mjs — pin the oracle to the REFERENCE ITSELF the first time a box has Node.js.
//
// Runs the UNMODIFIED worker body of vxlerian/cpu-raymarcher (src/workers/raymarchWorker.ts: `self.onmessage`) on every case
// of tests/golden/manifest.json and writes what it posts back — the reference's own Result planes — as
// tests/golden/ref/<case>.refbin (+ ref_meta.json).  tests/test_reference_fixtures.py then compares the oracle (and through
// it the CUDA validation build) with these files bit for bit: that turns "parity unpinned" into a pinned oracle.
//
// This container has no JS engine and no network (gl-matrix is an un-vendored npm dependency of the reference,
// package-lock.json:1258-1263), so the script cannot run here; it is one command elsewhere:
//
//     git clone https://github.com/vxlerian/cpu-raymarcher ref && (cd ref && npm ci)       # installs gl-matrix 3.4.4
//     npx tsx tools/ref_fixtures.mjs --reference ./ref [--out tests/golden/ref]
//
// Nothing of the reference is modified or copied: the worker module is imported as it is; `self` is the only thing provided.
// Synthetic-sphere cases (config 4: the reference has no large scene) use the reference's own classes the same way the
// worker does (`new Scene`, `new Sphere(worldToLocal, r)`, `alg.runRaymarcher(...)`) with spheres from the mulberry32
// generator of SURVEY.md §8d.  Layout of a .refbin: u8 depth[W*H] | u8 normal[3*W*H] | u16 sdfEval[W*H] | u16 iters[W*H] (LE).
import fs from 'node:fs';
import path from 'node:path';
import { pathToFileURL } from 'node:url';

const args = process.argv.slice(2);
const opt = (name, dflt) => { const i = args.indexOf(name); return i >= 0 ? args[i + 1] : dflt; };
const here = path.dirname(new URL(import.meta.url).pathname);
const ref = path.resolve(opt('--reference', process.env.RM_REFERENCE || './ref'));
const outDir = path.resolve(opt('--out', path.join(here, '..', 'tests', 'golden', 'ref')));
const manifest = JSON.parse(fs.readFileSync(path.join(here, '..', 'tests', 'golden', 'manifest.json'), 'utf8'));
fs.mkdirSync(outDir, { recursive: true });

// the worker global: capture exactly one reply per request (raymarchWorker.ts:86-91)
let reply = null;
globalThis.self = { onmessage: null, postMessage: (msg) => { reply = msg; } };
const imp = (p) => import(pathToFileURL(path.join(ref, p)).href);
await imp('src/workers/raymarchWorker.ts');            // sets self.onmessage; nothing else
const { Scene } = await imp('src/util/scene.ts');
const { Sphere } = await imp('src/util/primitives/sphere.ts');
const { SphereTracer } = await imp('src/cpu_algorithms/sphereTracer.ts');
const glm = await import(pathToFileURL(path.join(ref, 'node_modules', 'gl-matrix', 'esm', 'index.js')).href);
const { mat4, vec3 } = glm;

function mulberry32(seed) {                             // SURVEY.md §8d, 32-bit integer ops only
  let a = seed | 0;
  return () => {
    a = (a + 0x6D2B79F5) | 0;
    let t = Math.imul(a ^ (a >>> 15), 1 | a);
    t = (t + Math.imul(t ^ (t >>> 7), 61 | t)) ^ t;
    return ((t ^ (t >>> 14)) >>> 0) / 4294967296;
  };
}
function sphereAt(x, y, z, r) {                         // SceneManager.createSphere / getTransform (sceneManager.ts:21-41), no rotation
  const model = mat4.create();
  mat4.fromRotationTranslationScale(model, [0, 0, 0, 1], [x, y, z], [1, 1, 1]);
  const w2l = mat4.create();
  mat4.invert(w2l, model);
  return new Sphere(w2l, r);
}

function runSynthetic(c) {
  const [n, seed] = c.synthetic;
  const scene = new Scene('None');
  scene.loadPreset(1);                                  // "Random Spheres": prims 0-6 are the reference's own
  const objs = scene.objectSDFs.slice(0, Math.min(7, n));
  const rnd = mulberry32(seed);
  for (let i = 7; i < n; i++) {
    const cx = -2.5 + 5 * rnd(), cy = -2.5 + 5 * rnd(), cz = -2.5 + 5 * rnd(), r = 0.02 + 0.03 * rnd();
    objs.push(sphereAt(cx, cy, cz, r));
  }
  scene.accelerationStructure = c.accel;
  scene.objectSDFs = objs;
  // rebuild the acceleration structure over the new object list through the reference's own (TS-private) builders
  if (c.accel === 'BVH') { scene.cachedBounds = null; scene.buildBVH(); scene.octree = null; }
  else if (c.accel === 'Octree') { scene.cachedBounds = null; scene.buildOctree(); scene.bvh = null; }
  scene.camera.setAngles(c.pitch, c.yaw);
  const n_px = c.W * c.H;
  const out = { depth: new Uint8ClampedArray(n_px), normal: new Uint8ClampedArray(3 * n_px), sdfEval: new Uint16Array(n_px), iters: new Uint16Array(n_px) };
  if (c.alg !== 'sphere-tracer') throw new Error('synthetic cases use the sphere tracer');
  new SphereTracer().runRaymarcher(scene, out.depth, out.normal, out.sdfEval, out.iters, c.W, c.H, c.time, 0, c.H);
  return out;
}

const meta = {
  generator: 'tools/ref_fixtures.mjs', reference: ref, node: process.version,
  gl_matrix_version: JSON.parse(fs.readFileSync(path.join(ref, 'node_modules', 'gl-matrix', 'package.json'), 'utf8')).version,
  // the open question of SURVEY.md Appendix B: does vec3.length call Math.hypot (V8 Kahan) or Math.sqrt(x*x+y*y+z*z)?
  vec3_length_source: String(vec3.length), vec3_length_uses_hypot: /Math\.hypot/.test(String(vec3.length)),
  cases: [],
};
for (const c of manifest.cases) {
  let r;
  if (c.synthetic) {
    r = runSynthetic(c);
  } else {
    reply = null;
    self.onmessage({ data: { width: c.W, height: c.H, time: c.time, yStart: 0, yEnd: c.H, camera: { pitch: c.pitch, yaw: c.yaw }, algorithm: c.alg,
                             scenePresetIndex: c.preset, accelerationStructure: c.accel, overshootFactor: c.over, stepSize: c.step } });
    if (!reply) throw new Error(`${c.name}: the worker posted no reply`);
    r = reply;
  }
  const buf = Buffer.concat([Buffer.from(r.depth.buffer, r.depth.byteOffset, r.depth.byteLength), Buffer.from(r.normal.buffer, r.normal.byteOffset, r.normal.byteLength),
                             Buffer.from(r.sdfEval.buffer, r.sdfEval.byteOffset, r.sdfEval.byteLength), Buffer.from(r.iters.buffer, r.iters.byteOffset, r.iters.byteLength)]);
  fs.writeFileSync(path.join(outDir, c.name + '.refbin'), buf);
  meta.cases.push({ name: c.name, bytes: buf.length });
  console.log(`${c.name}: ${buf.length} bytes`);
}
fs.writeFileSync(path.join(outDir, 'ref_meta.json'), JSON.stringify(meta, null, 1));
console.log(`vec3.length uses Math.hypot: ${meta.vec3_length_uses_hypot}  (gl-matrix ${meta.gl_matrix_version})`);
