// ORACLE — TEST INFRASTRUCTURE ONLY (see jsnum.hpp header).  PARITY UNPINNED.
//
// march.hpp — the per-tile pixel loop, the five march loops, the four shaders, the stats loop.
// Follows, in order:
//   src/cpu_algorithms/raymarcher.ts:46-135   runRaymarcher / getSceneDistance / getNormal
//   src/cpu_algorithms/sphereTracer.ts:15-83
//   src/cpu_algorithms/fixedStep.ts:21-94
//   src/cpu_algorithms/adaptiveStep.ts:22-105
//   src/cpu_algorithms/adaptiveStepV2.ts:22-124
//   src/cpu_algorithms/adaptiveStepV3.ts:22-137
//   src/util/shading_models/{normalModel,phongModel,SDFHeatmap,IterationHeatmap}.ts
//   src/main.ts:527-548                       diagnostics loop
#pragma once
#include "scene.hpp"

namespace orc {

enum Algo { ALG_SPHERE = 0, ALG_FIXED = 1, ALG_ADAPTIVE = 2, ALG_V2 = 3, ALG_V3 = 4 };
enum Shader { SH_NORMAL = 0, SH_PHONG = 1, SH_SDF_HEAT = 2, SH_ITER_HEAT = 3 };

struct MarchParams {
    int algo = ALG_SPHERE;
    double stepSize = 0.1;         // fixedStep.ts:13
    double overshootFactor = 1.2;  // adaptiveStepV2.ts:13
};

// Per-pixel outputs.  u16 fields wrap like the reference's Uint16Arrays; the *_full counters are the
// un-wrapped totals (used only for throughput accounting, never for parity of the reference contract).
struct PixelOut {
    uint16_t sdfEval = 0, iters = 0;
    uint64_t sdfEvalFull = 0, itersFull = 0;
    double depth = 0;  // unquantised return of rayMarch
};

struct Raymarcher {
    const Scene& scene;
    MarchParams mp;
    Raymarcher(const Scene& s, const MarchParams& p) : scene(s), mp(p) {}

    static constexpr double MAX_DIST = 10;
    static constexpr double EPSILON = 0.001;

    // raymarcher.ts:111-121
    double getSceneDistance(const vec3& position, PixelOut& px) const {
        uint32_t count = 0;
        double d = scene.getDistance(position, count);
        px.sdfEval = (uint16_t)(px.sdfEval + count);  // Uint16Array += : wraps mod 65536
        px.sdfEvalFull += count;
        return d;
    }
    // raymarcher.ts:123-135
    vec3 getNormal(const vec3& position, PixelOut& px) const {
        double d = getSceneDistance(position, px);
        const double e0 = 0.01;
        vec3 n = glm::v3_create();
        n.e[0] = js::f32(d - getSceneDistance(glm::v3_from(position[0] - e0, position[1], position[2]), px));
        n.e[1] = js::f32(d - getSceneDistance(glm::v3_from(position[0], position[1] - e0, position[2]), px));
        n.e[2] = js::f32(d - getSceneDistance(glm::v3_from(position[0], position[1], position[2] - e0), px));
        glm::v3_normalize(n, n);
        return n;
    }

    // Acceleration-structure callback protocol (accelerationStructure.ts:23-39), resolved statically.
    struct AccelState {
        bool present = false;  // accelState truthy
        BVHState bvh;
    };
    // returns false when rayMarch must `return MAX_DIST` immediately (BVH found nothing)
    bool accelStart(const vec3& o, const vec3& d, AccelState& st) const {
        if (scene.accel == ACCEL_BVH && scene.bvh) {
            st.present = true;
            st.bvh = scene.bvh->onRayMarchStart(o, d, MAX_DIST);
            if (st.bvh.terminate) return false;
        } else if (scene.accel == ACCEL_OCTREE && scene.octree) {
            st.present = true;  // {data:null}
        }
        return true;
    }
    double accelStep(const vec3& o, const vec3& d, double t, AccelState& st) const {
        if (!st.present) return 0;
        if (scene.accel == ACCEL_BVH) return scene.bvh->onRayMarchStep(t, st.bvh);
        return scene.octree->marchRay(o, d, t);
    }

    inline void bumpIters(PixelOut& px) const {
        px.iters = (uint16_t)(px.iters + 1);
        px.itersFull += 1;
    }

    double rayMarch(const vec3& o, const vec3& dir, PixelOut& px) const {
        switch (mp.algo) {
            case ALG_FIXED: return marchFixed(o, dir, px);
            case ALG_ADAPTIVE: return marchAdaptive(o, dir, px);
            case ALG_V2: return marchV2(o, dir, px);
            case ALG_V3: return marchV3(o, dir, px);
            default: return marchSphere(o, dir, px);
        }
    }

    // sphereTracer.ts:15-83
    double marchSphere(const vec3& o, const vec3& dir, PixelOut& px) const {
        const int MAX_STEPS = 100;
        double totalDist = 0;
        AccelState st;
        if (!accelStart(o, dir, st)) return MAX_DIST;
        for (int i = 0; i < MAX_STEPS; ++i) {
            vec3 p = glm::v3_create();
            glm::v3_scale_and_add(p, o, dir, totalDist);
            if (st.present) {
                double skip = accelStep(o, dir, totalDist, st);
                if (skip == -1) return MAX_DIST;
                else if (skip > 0) {
                    totalDist += skip;
                    if (totalDist > MAX_DIST) break;
                    continue;
                }
            }
            double dist = getSceneDistance(p, px);
            totalDist += dist;
            bumpIters(px);
            if (dist < EPSILON) break;
            if (totalDist > MAX_DIST) break;
        }
        return totalDist;
    }
    // fixedStep.ts:21-94
    double marchFixed(const vec3& o, const vec3& dir, PixelOut& px) const {
        const int MAX_STEPS = 200;
        double totalDist = 0;
        bool hit = false;
        AccelState st;
        if (!accelStart(o, dir, st)) return MAX_DIST;
        for (int i = 0; i < MAX_STEPS; ++i) {
            vec3 p = glm::v3_create();
            glm::v3_scale_and_add(p, o, dir, totalDist);
            if (st.present) {
                double skip = accelStep(o, dir, totalDist, st);
                if (skip == -1) return MAX_DIST;
                else if (skip > 0) {
                    totalDist += skip;
                    if (totalDist > MAX_DIST) break;
                    continue;
                }
            }
            double dist = getSceneDistance(p, px);
            bumpIters(px);
            if (dist < EPSILON) {
                hit = true;
                break;
            }
            totalDist += mp.stepSize;
            if (totalDist > MAX_DIST) break;
        }
        return hit ? totalDist : MAX_DIST;
    }
    // adaptiveStep.ts:22-105
    double marchAdaptive(const vec3& o, const vec3& dir, PixelOut& px) const {
        const int MAX_STEPS = 200;
        const double FIXED_STEP_SIZE = 0.1;
        const double STEP_SCALE = 0.8;
        const double MIN_STEP = FIXED_STEP_SIZE * 0.25;
        const double MAX_STEP = FIXED_STEP_SIZE * 5.0;
        const double NEAR_DIST = 0.1;
        const double NEAR_STEP = 0.01;
        double totalDist = 0;
        bool hit = false;
        AccelState st;
        if (!accelStart(o, dir, st)) return MAX_DIST;
        for (int i = 0; i < MAX_STEPS; ++i) {
            vec3 p = glm::v3_create();
            glm::v3_scale_and_add(p, o, dir, totalDist);
            if (st.present) {
                double skip = accelStep(o, dir, totalDist, st);
                if (skip == -1) return MAX_DIST;
                else if (skip > 0) {
                    totalDist += skip;
                    if (totalDist > MAX_DIST) break;
                    continue;
                }
            }
            double dist = getSceneDistance(p, px);
            bumpIters(px);
            if (dist < EPSILON) {
                hit = true;
                break;
            }
            double step;
            if (dist < NEAR_DIST) {
                step = NEAR_STEP;
            } else {
                step = STEP_SCALE * dist;
                if (step < MIN_STEP) step = MIN_STEP;
                if (step > MAX_STEP) step = MAX_STEP;
            }
            totalDist += step;
            if (totalDist > MAX_DIST) break;
        }
        return hit ? totalDist : MAX_DIST;
    }
    // adaptiveStepV2.ts:22-124
    double marchV2(const vec3& o, const vec3& dir, PixelOut& px) const {
        const int MAX_STEPS = 100;
        double totalDist = 0, prevSDF = 0, prevStep = 0;
        AccelState st;
        if (!accelStart(o, dir, st)) return MAX_DIST;
        for (int i = 0; i < MAX_STEPS; ++i) {
            vec3 p = glm::v3_create();
            glm::v3_scale_and_add(p, o, dir, totalDist);
            if (st.present) {
                double skip = accelStep(o, dir, totalDist, st);
                if (skip == -1) return MAX_DIST;
                else if (skip > 0) {
                    totalDist += skip;
                    if (totalDist > MAX_DIST) break;
                    prevSDF = 0;
                    prevStep = 0;
                    continue;
                }
            }
            double newSDF = getSceneDistance(p, px);
            bumpIters(px);
            if (newSDF < EPSILON) break;
            if (totalDist > MAX_DIST) break;
            if (i == 0 || prevSDF == 0) {
                double step = newSDF;
                totalDist += step;
                prevSDF = newSDF;
                prevStep = step;
            } else {
                bool spheresOverlapped = prevStep <= (prevSDF + newSDF);
                if (spheresOverlapped) {
                    double step = newSDF * mp.overshootFactor;
                    totalDist += step;
                    prevSDF = newSDF;
                    prevStep = step;
                } else {
                    totalDist -= prevStep;
                    totalDist += prevSDF;
                    prevStep = prevSDF;
                }
            }
        }
        return totalDist;
    }
    // adaptiveStepV3.ts:22-137
    double marchV3(const vec3& o, const vec3& dir, PixelOut& px) const {
        const int MAX_STEPS = 100;
        double totalDist = 0, prevSDF = 0, prevStep = 0;
        AccelState st;
        if (!accelStart(o, dir, st)) return MAX_DIST;
        vec3 tmpP = glm::v3_create();
        for (int i = 0; i < MAX_STEPS; ++i) {
            glm::v3_scale_and_add(tmpP, o, dir, totalDist);
            if (st.present) {
                double skip = accelStep(o, dir, totalDist, st);
                if (skip == -1) return MAX_DIST;
                else if (skip > 0) {
                    totalDist += skip;
                    if (totalDist > MAX_DIST) break;
                    prevSDF = 0;
                    prevStep = 0;
                    continue;
                }
            }
            double newSDF = getSceneDistance(tmpP, px);
            bumpIters(px);
            if (newSDF < EPSILON) break;
            if (totalDist > MAX_DIST) break;
            if (i == 0 || prevSDF == 0) {
                double step = newSDF;
                totalDist += step;
                prevSDF = newSDF;
                prevStep = step;
                continue;
            }
            bool spheresOverlapped = prevStep <= (prevSDF + newSDF);
            if (spheresOverlapped) {
                double step = newSDF * mp.overshootFactor;
                totalDist += step;
                prevSDF = newSDF;
                prevStep = step;
                continue;
            }
            double originalPos = totalDist - prevStep;
            totalDist = originalPos + prevSDF;
            glm::v3_scale_and_add(tmpP, o, dir, totalDist);
            double d3 = getSceneDistance(tmpP, px);
            bumpIters(px);
            if (prevSDF + newSDF + d3 >= prevStep) {
                totalDist = originalPos + prevStep + newSDF;
                prevSDF = newSDF;
                prevStep = newSDF;
                continue;
            }
            prevSDF = d3;
            prevStep = d3;
            totalDist += d3;
        }
        return totalDist;
    }

    // raymarcher.ts:46-109.  Buffers are tile-local ((y - yStart) * width + x).
    // Optional extras (may be null): depthF64 (unquantised rayMarch return), sdfFull/itersFull (unwrapped).
    void runRows(int width, int height, int yStart, int yEnd, int rowBegin, int rowStride, uint8_t* depthBuffer,
                 uint8_t* normalBuffer, uint16_t* sdfBuffer, uint16_t* itersBuffer, double* depthF64,
                 uint32_t* sdfFull, uint32_t* itersFull) const {
        for (int y = yStart + rowBegin; y < yEnd; y += rowStride)
            runRow(width, height, y, y - yStart, depthBuffer, normalBuffer, sdfBuffer, itersBuffer, depthF64, sdfFull, itersFull);
    }
    // one image row y, written at buffer row localY
    void runRow(int width, int height, int y, int localY, uint8_t* depthBuffer, uint8_t* normalBuffer, uint16_t* sdfBuffer,
                uint16_t* itersBuffer, double* depthF64, uint32_t* sdfFull, uint32_t* itersFull) const {
        mat4 rotMat4;
        scene.camera.getRotationMatrix(rotMat4);
        mat3 rotMat3;
        glm::m3_from_mat4(rotMat3, rotMat4);
        vec3 rayOrigin = glm::v3_create();
        scene.camera.getPosition(rayOrigin);
        {
            double v = ((double)y / (double)height - 0.5) * 2.0;
            for (int x = 0; x < width; ++x) {
                size_t idx = (size_t)localY * width + x;
                size_t normalIdx = idx * 3;
                PixelOut px;
                double u = ((double)x / (double)width - 0.5) * 2.0;
                vec3 rayDir = glm::v3_from(u, v, -1);
                glm::v3_transform_mat3(rayDir, rayDir, rotMat3);
                glm::v3_normalize(rayDir, rayDir);
                double depth = rayMarch(rayOrigin, rayDir, px);
                px.depth = depth;
                vec3 hitPosition = glm::v3_create();
                glm::v3_scale_and_add(hitPosition, rayOrigin, rayDir, depth);
                vec3 normal;
                if (depth >= MAX_DIST) normal = glm::v3_from(0, 0, 0);
                else normal = getNormal(hitPosition, px);
                normalBuffer[normalIdx] = js::to_u8_clamp((normal[0] + 1) * 0.5 * 255);
                normalBuffer[normalIdx + 1] = js::to_u8_clamp((normal[1] + 1) * 0.5 * 255);
                normalBuffer[normalIdx + 2] = js::to_u8_clamp((normal[2] + 1) * 0.5 * 255);
                depthBuffer[idx] = js::to_u8_clamp(depth);
                sdfBuffer[idx] = px.sdfEval;
                itersBuffer[idx] = px.iters;
                if (depthF64) depthF64[idx] = depth;
                if (sdfFull) sdfFull[idx] = (uint32_t)px.sdfEvalFull;
                if (itersFull) itersFull[idx] = (uint32_t)px.itersFull;
            }
        }
    }
};

// ------------------------------------------------------------------ shading models
inline void shadeHeat(uint8_t* out, const uint16_t* counts, size_t n) {  // SDFHeatmap.ts:19-31 / IterationHeatmap.ts:19-31
    const double colourScalingFactor = 5;
    for (size_t idx = 0; idx < n; ++idx) {
        double sdfIntensity = std::fmod((double)counts[idx] * colourScalingFactor, 256.0);
        out[idx * 4 + 0] = js::to_u8_clamp(js::min2(2 * sdfIntensity, 255));
        out[idx * 4 + 1] = js::to_u8_clamp(js::min2(-2 * sdfIntensity + 512, 255));
        out[idx * 4 + 2] = 0;
        out[idx * 4 + 3] = 255;
    }
}
inline void shadeNormal(uint8_t* out, const uint8_t* normal, size_t n) {  // normalModel.ts:15-26
    for (size_t idx = 0; idx < n; ++idx) {
        out[idx * 4 + 0] = normal[idx * 3 + 0];
        out[idx * 4 + 1] = normal[idx * 3 + 1];
        out[idx * 4 + 2] = normal[idx * 3 + 2];
        out[idx * 4 + 3] = 255;
    }
}
inline void shadePhong(uint8_t* out, const uint8_t* depthBuffer, const uint8_t* normalBuffer, size_t n) {  // phongModel.ts:15-73
    vec3 lightDir = glm::v3_from(1, -1, 1.5);
    glm::v3_normalize(lightDir, lightDir);
    vec3 normal = glm::v3_create(), reflectDir = glm::v3_create();
    vec3 viewDir = glm::v3_from(0, 0, 1);
    const double ambient = 0.1, specularStrength = 0.5, shininess = 32;
    for (size_t idx = 0; idx < n; ++idx) {
        double depth = depthBuffer[idx];
        if (depth >= 255) {
            out[idx * 4 + 0] = 10;
            out[idx * 4 + 1] = 10;
            out[idx * 4 + 2] = 20;
            out[idx * 4 + 3] = 255;
            continue;
        }
        normal.e[0] = js::f32(normalBuffer[idx * 3 + 0] / 127.5 - 1.0);
        normal.e[1] = js::f32(normalBuffer[idx * 3 + 1] / 127.5 - 1.0);
        normal.e[2] = js::f32(normalBuffer[idx * 3 + 2] / 127.5 - 1.0);
        glm::v3_normalize(normal, normal);
        double diffuse = js::max2(glm::v3_dot(normal, lightDir), 0);
        glm::v3_scale(reflectDir, normal, 2 * glm::v3_dot(normal, lightDir));
        glm::v3_subtract(reflectDir, reflectDir, lightDir);
        glm::v3_normalize(reflectDir, reflectDir);
        double specular = specularStrength * std::pow(js::max2(glm::v3_dot(viewDir, reflectDir), 0), shininess);
        double intensity = js::min2(ambient + diffuse + specular, 1);
        double depthFactor = 1 - depth / 255;
        double color = 255 * intensity * depthFactor;
        uint8_t c = js::to_u8_clamp(color);
        out[idx * 4 + 0] = c;
        out[idx * 4 + 1] = c;
        out[idx * 4 + 2] = c;
        out[idx * 4 + 3] = 255;
    }
}

// main.ts:527-548 (on the wrapped u16 buffers)
struct FrameStats {
    double totalSDFCalls = 0, maxSDFCalls = 0, minSDFCalls = 9007199254740991.0, totalIterations = 0;
};
inline FrameStats frameStats(const uint16_t* sdf, const uint16_t* iters, size_t n) {
    FrameStats s;
    for (size_t i = 0; i < n; ++i) {
        double c = sdf[i];
        s.totalSDFCalls += c;
        s.totalIterations += iters[i];
        if (c > s.maxSDFCalls) s.maxSDFCalls = c;
        if (c < s.minSDFCalls) s.minSDFCalls = c;
    }
    return s;
}

}  // namespace orc
