// Type declarations of the rm_napi addon (addon/rm_napi.cc): what ts/gpuWorkerShim.ts calls.
// Job / Result are the reference's worker contract (src/workers/raymarchWorker.ts:10-31).

export interface SceneArrays {
  /** rm_prim_type per primitive: 0 sphere, 1 box, 2 torus, 3 mandelbulb */
  type: Uint8Array;
  /** Primitive.transform of every primitive, column-major mat4 (16 floats each) */
  worldToLocal: Float32Array;
  /** 4 doubles per primitive: sphere r | box hx,hy,hz | torus R,r | mandelbulb power,iterations,enableAnimation,animationSpeed */
  params: Float64Array;
  accel: 'None' | 'Octree' | 'BVH';
  /** operator trees: rm_op_node records, 128 bytes each (include/rm.h) */
  opNodes?: Uint8Array;
  /** root node index of every scene object (present iff opNodes is) */
  objectRoot?: Int32Array;
}

export interface Job {
  width: number; height: number; time: number; yStart: number; yEnd: number;
  algorithm: string; overshootFactor?: number; stepSize?: number;
}

export interface CameraBasis {
  /** mat3.fromMat4(camera.getRotationMatrix()) — raymarcher.ts:62-64 */
  rot3: Float32Array;
  /** camera.getPosition() — raymarcher.ts:66-67 */
  origin: Float32Array;
}

export interface Result {
  yStart: number; yEnd: number;
  depth: Uint8ClampedArray; normal: Uint8ClampedArray; sdfEval: Uint16Array; iters: Uint16Array;
}

export interface FrameStats {
  totalSDFCalls: number; maxSDFCalls: number; minSDFCalls: number; totalIterations: number;
  totalPixels: number; kernelMs: number; devices: number;
}

/** rm_pool_upload_scene: compiled and uploaded once, replicated to every GPU of the pool. Throws on error. */
export function uploadScene(scene: SceneArrays): void;
/** rm_pool_render on a libuv worker thread. Band Jobs of one frame share a single render across all GPUs. Rejects on error. */
export function render(job: Job, camera: CameraBasis): Promise<Result>;
/** rm_pool_stats: diagnostics of main.ts:527-548 over the whole frame of the last Job. */
export function stats(): FrameStats;
/** GPUs owned by the pool (all visible ones, or RM_DEVICES=0,1,...). */
export function deviceCount(): number;
