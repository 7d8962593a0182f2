// ORACLE — TEST INFRASTRUCTURE ONLY (see jsnum.hpp header).  PARITY UNPINNED.
//
// scene.hpp — primitives, bounding boxes, BVH, octree, camera, presets, Scene.getDistance.
// Follows, in order:
//   src/util/primitives/primitive.ts:20-39      Primitive.getWorldPosition / sdf
//   src/util/primitives/sphere.ts:12-17         Sphere.localSdf / bounding radius
//   src/util/primitives/box.ts:13-34            Box.localSdf / bounding radius
//   src/util/primitives/torus.ts:14-29          Torus.localSdf / bounding radius
//   src/acceleration_structures/boundingBox.ts  whole file
//   src/acceleration_structures/bvh.ts          whole file
//   src/acceleration_structures/octree.ts       whole file
//   src/util/camera.ts:38-44,58-69,81-88        Camera
//   src/util/sceneManager.ts:21-49,73-100,102-356 getTransform / create* / presets 0-12,14-18
//   src/util/primitive_operations/*.ts          Round, Twist, SmoothUnion, SmoothSubtraction, Repetition, AnimatedTranslate
//   src/util/scene.ts:24-85,144-190             Scene
#pragma once
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "glm.hpp"

namespace orc {
using glm::mat3;
using glm::mat4;
using glm::vec3;

// Leaf kinds, then the operator kinds of src/util/primitive_operations/*.ts.
enum PrimType {
    SPHERE = 0,
    BOX = 1,
    TORUS = 2,
    ROUND = 3,               // round.ts
    TWIST = 4,               // twist.ts
    SMOOTH_UNION = 5,        // smoothUnion.ts
    SMOOTH_SUBTRACTION = 6,  // smoothSubstraction.ts
    REPETITION = 7,          // repetition.ts
    ANIMATED_TRANSLATE = 8,  // animatedTranslate.ts
    MANDELBULB = 9           // primitives/mandelbulb.ts (a leaf; rm_prim_type 3 in the flat arrays)
};

// primitive.ts:3-44 + sphere.ts / box.ts / torus.ts + the six operator classes.  One tagged struct instead
// of a class hierarchy; operators hold their wrapped primitives in `a` (primitive / prim1) and `b` (prim2).
struct Primitive {
    int type = SPHERE;
    int index = 0;    // position in Scene.objectSDFs (identity, used when flattening)
    mat4 transform;   // world -> local (primitive.ts:4).  Operators: super(primitive.transform) or identity.
    double radius = 0;                       // sphere.ts:5 (JS number = f64) | round.ts:6 radius
    vec3 halfSize = glm::v3_create();        // box.ts:10 (vec3.clone -> f32)
    double majorRadius = 0, minorRadius = 0;  // torus.ts:5-6 (f64)
    std::shared_ptr<Primitive> a, b;          // wrapped primitive(s)
    double twistAmount = 0;                   // twist.ts:6
    double smoothness = 0;                    // smoothUnion.ts:7 / smoothSubstraction.ts:7
    vec3 spacing = glm::v3_create();          // repetition.ts:6 (the caller's vec3, f32)
    vec3 direction = glm::v3_create();        // animatedTranslate.ts:9 (normalised, f32)
    double amplitude = 0, speed = 0, time = 0;  // animatedTranslate.ts:10-12 (time also: mandelbulb.ts:16)
    double power = 8.0;                         // mandelbulb.ts:12-15
    int iterations = 9;
    bool enableAnimation = true;
    double animationSpeed = -0.2;

    // primitive.ts:20-30 and the operator overrides
    vec3 getWorldPosition() const {
        switch (type) {
            case ROUND: case TWIST: case REPETITION: case ANIMATED_TRANSLATE:  // round.ts:31-34, twist.ts:43-46, ...
            case SMOOTH_SUBTRACTION:                                            // smoothSubstraction.ts:41-44
                return a->getWorldPosition();
            case SMOOTH_UNION: {  // smoothUnion.ts:50-59
                vec3 pos1 = a->getWorldPosition(), pos2 = b->getWorldPosition();
                return glm::v3_from((pos1[0] + pos2[0]) / 2, (pos1[1] + pos2[1]) / 2, (pos1[2] + pos2[2]) / 2);
            }
            default: break;
        }
        mat4 localToWorld = glm::m4_create();
        glm::m4_invert(localToWorld, transform);
        return glm::v3_from(localToWorld[12], localToWorld[13], localToWorld[14]);
    }
    // sphere.ts:16, box.ts:32-34, torus.ts:27-29 and the operator overrides
    double getLocalBoundingRadius() const {
        switch (type) {
            case SPHERE: return radius;
            case BOX: return glm::v3_length(halfSize);
            case TORUS: return majorRadius + minorRadius;
            case ROUND: return a->getLocalBoundingRadius() + radius;              // round.ts:26-29
            case TWIST: return a->getLocalBoundingRadius();                       // twist.ts:38-41
            case SMOOTH_SUBTRACTION: return a->getLocalBoundingRadius();          // smoothSubstraction.ts:36-39
            case REPETITION: return js::kInf;                                     // repetition.ts:31-34
            case MANDELBULB: return 2.5;                                          // mandelbulb.ts:80-83
            case ANIMATED_TRANSLATE: return a->getLocalBoundingRadius() + amplitude;  // animatedTranslate.ts:50-53
            default: {                                                            // smoothUnion.ts:37-48
                double r1 = a->getLocalBoundingRadius(), r2 = b->getLocalBoundingRadius();
                vec3 pos1 = a->getWorldPosition(), pos2 = b->getWorldPosition();
                double centerDist = glm::v3_distance(pos1, pos2);
                return js::max2(r1, r2) + centerDist * 0.5;
            }
        }
    }
    // The operators' common prologue: "convert local position back to world space" (round.ts:17-20 etc.).
    vec3 backToWorld(const vec3& localPos) const {
        vec3 worldPos = glm::v3_create();
        mat4 localToWorld = glm::m4_create();
        glm::m4_invert(localToWorld, transform);
        glm::v3_transform_mat4(worldPos, localPos, localToWorld);
        return worldPos;
    }
    double localSdf(const vec3& l) const {
        switch (type) {
            case SPHERE:  // sphere.ts:12-14
                return glm::v3_length(l) - radius;
            case BOX: {  // box.ts:13-30
                vec3 q = glm::v3_from(std::fabs(l[0]) - halfSize[0], std::fabs(l[1]) - halfSize[1],
                                      std::fabs(l[2]) - halfSize[2]);
                vec3 outside = glm::v3_from(js::max2(q[0], 0), js::max2(q[1], 0), js::max2(q[2], 0));
                double outsideDist = glm::v3_length(outside);
                double insideDist = js::min2(js::max2(q[0], js::max2(q[1], q[2])), 0);
                return outsideDist + insideDist;
            }
            case TORUS: {  // torus.ts:14-25
                double x = l[0], y = l[1], z = l[2];
                double qx = std::sqrt(x * x + z * z) - majorRadius;
                double qy = y;
                return std::sqrt(qx * qx + qy * qy) - minorRadius;
            }
            case MANDELBULB: {  // mandelbulb.ts:38-78
                vec3 p = glm::v3_from(l[0], l[2], l[1]);  // p.xyz = p.xzy
                vec3 z = p;
                double dr = 1.0, r = 0.0;
                for (int i = 0; i < iterations; ++i) {
                    r = glm::v3_length(z);
                    if (r > 2.0) break;
                    double theta = js::atan2(z[1], z[0]);
                    double phi = js::asin(z[2] / r);
                    if (enableAnimation) phi += time * animationSpeed;
                    dr = js::pow(r, power - 1.0) * dr * power + 1.0;
                    r = js::pow(r, power);  // NB: the loop variable r now holds r^power (it is what the return uses after the last pass)
                    theta = theta * power;
                    phi = phi * power;
                    z.e[0] = js::f32(r * js::cos(theta) * js::cos(phi) + p[0]);
                    z.e[1] = js::f32(r * js::sin(theta) * js::cos(phi) + p[1]);
                    z.e[2] = js::f32(r * js::sin(phi) + p[2]);
                }
                return 0.5 * js::log(r) * r / dr;
            }
            case ROUND:  // round.ts:15-24
                return a->sdf(backToWorld(l)) - radius;
            case TWIST: {  // twist.ts:14-36
                vec3 worldPos = backToWorld(l);
                const double k = twistAmount;
                const double c = js::cos(k * worldPos[1]);
                const double s = js::sin(k * worldPos[1]);
                vec3 twistedPos = glm::v3_from(c * worldPos[0] - s * worldPos[2], worldPos[1], s * worldPos[0] + c * worldPos[2]);
                return a->sdf(twistedPos);
            }
            case SMOOTH_UNION: {  // smoothUnion.ts:18-35
                vec3 worldPos = backToWorld(l);
                const double d1 = a->sdf(worldPos), d2 = b->sdf(worldPos);
                const double k = smoothness * 4.0;
                const double h = js::max2(k - std::fabs(d1 - d2), 0.0);
                return js::min2(d1, d2) - h * h * 0.25 / k;
            }
            case SMOOTH_SUBTRACTION: {  // smoothSubstraction.ts:17-34
                vec3 worldPos = backToWorld(l);
                const double d1 = a->sdf(worldPos), d2 = b->sdf(worldPos);
                const double k = smoothness * 4.0;
                const double h = js::max2(k - std::fabs(d1 + d2), 0.0);
                return js::max2(d1, -d2) + h * h * 0.25 / k;
            }
            case REPETITION: {  // repetition.ts:14-29
                vec3 worldPos = backToWorld(l);
                vec3 q = glm::v3_create();
                q.e[0] = js::f32(worldPos[0] - spacing[0] * js::round(worldPos[0] / spacing[0]));
                q.e[1] = js::f32(worldPos[1] - spacing[1] * js::round(worldPos[1] / spacing[1]));
                q.e[2] = js::f32(worldPos[2] - spacing[2] * js::round(worldPos[2] / spacing[2]));
                return a->sdf(q);
            }
            default: {  // animatedTranslate.ts:34-48 (no back-to-world step: the child re-applies the transform)
                const double offset = js::sin(time * speed) * amplitude;
                vec3 offsetVec = glm::v3_create();
                glm::v3_scale(offsetVec, direction, offset);
                vec3 adjustedPos = glm::v3_create();
                glm::v3_subtract(adjustedPos, l, offsetVec);
                return a->sdf(adjustedPos);
            }
        }
    }
    // primitive.ts:33-39
    double sdf(const vec3& pos) const {
        vec3 localPos = glm::v3_create();
        glm::v3_transform_mat4(localPos, pos, transform);
        return localSdf(localPos);
    }
    // primitive.ts:42-44 and the overrides (animatedTranslate.ts:30-32 does NOT forward to its child)
    void setTime(double t) {
        switch (type) {
            case ANIMATED_TRANSLATE: case MANDELBULB: time = t; break;  // animatedTranslate.ts:30-32, mandelbulb.ts:33-35
            case ROUND: case TWIST: case REPETITION: a->setTime(t); break;
            case SMOOTH_UNION: case SMOOTH_SUBTRACTION: a->setTime(t); b->setTime(t); break;
            default: break;
        }
    }
};

// ------------------------------------------------------------------ boundingBox.ts
struct BoundingBox {
    vec3 min, max;
    BoundingBox() : min(glm::v3_create()), max(glm::v3_create()) {}
    BoundingBox(const vec3& mn, const vec3& mx) : min(mn), max(mx) {}

    bool contains(const vec3& p) const {  // :15-21
        return p[0] >= min[0] && p[0] <= max[0] && p[1] >= min[1] && p[1] <= max[1] && p[2] >= min[2] &&
               p[2] <= max[2];
    }
    bool intersects(const BoundingBox& o) const {  // :24-30
        return min[0] <= o.max[0] && max[0] >= o.min[0] && min[1] <= o.max[1] && max[1] >= o.min[1] &&
               min[2] <= o.max[2] && max[2] >= o.min[2];
    }
    double distanceToBox(const BoundingBox& o) const {  // :33-47
        double dx = 0;
        if (max[0] < o.min[0]) dx = o.min[0] - max[0];
        else if (o.max[0] < min[0]) dx = min[0] - o.max[0];
        double dy = 0;
        if (max[1] < o.min[1]) dy = o.min[1] - max[1];
        else if (o.max[1] < min[1]) dy = min[1] - o.max[1];
        double dz = 0;
        if (max[2] < o.min[2]) dz = o.min[2] - max[2];
        else if (o.max[2] < min[2]) dz = min[2] - o.max[2];
        return js::hypot3(dx, dy, dz);
    }
    // :69-105  returns false for null
    bool intersectRay(const vec3& origin, const vec3& direction, double& tMinOut, double& tMaxOut) const {
        double tMin = -js::kInf, tMax = js::kInf;
        for (int i = 0; i < 3; ++i) {
            if (std::fabs(direction[i]) < 1e-10) {
                if (origin[i] < min[i] || origin[i] > max[i]) return false;
            } else {
                double invD = 1.0 / direction[i];
                double t0 = (min[i] - origin[i]) * invD;
                double t1 = (max[i] - origin[i]) * invD;
                if (t0 > t1) std::swap(t0, t1);
                tMin = js::max2(tMin, t0);
                tMax = js::min2(tMax, t1);
                if (tMin > tMax) return false;
            }
        }
        tMinOut = tMin;
        tMaxOut = tMax;
        return true;
    }
    vec3 center() const {  // :108-114
        return glm::v3_from((min[0] + max[0]) / 2, (min[1] + max[1]) / 2, (min[2] + max[2]) / 2);
    }
    BoundingBox merge(const BoundingBox& o) const {  // :117-130
        return BoundingBox(
            glm::v3_from(js::min2(min[0], o.min[0]), js::min2(min[1], o.min[1]), js::min2(min[2], o.min[2])),
            glm::v3_from(js::max2(max[0], o.max[0]), js::max2(max[1], o.max[1]), js::max2(max[2], o.max[2])));
    }
    static BoundingBox fromPrimitive(const Primitive& p) {  // :133-154
        vec3 worldPos = p.getWorldPosition();
        double localRadius = p.getLocalBoundingRadius();
        mat4 localToWorld = glm::m4_create();
        bool ok = glm::m4_invert(localToWorld, p.transform);
        const mat4& m = ok ? localToWorld : p.transform;
        double scaleX = js::hypot3(m[0], m[1], m[2]);
        double scaleY = js::hypot3(m[4], m[5], m[6]);
        double scaleZ = js::hypot3(m[8], m[9], m[10]);
        double maxScale = js::max3(scaleX, scaleY, scaleZ);
        double r = localRadius * maxScale * 1.5;
        return BoundingBox(glm::v3_from(worldPos[0] - r, worldPos[1] - r, worldPos[2] - r),
                           glm::v3_from(worldPos[0] + r, worldPos[1] + r, worldPos[2] + r));
    }
};

using PrimList = std::vector<const Primitive*>;

inline BoundingBox computeBounds(const PrimList& prims) {  // boundingBox.ts:158-169
    if (prims.empty()) return BoundingBox(glm::v3_from(0, 0, 0), glm::v3_from(0, 0, 0));
    BoundingBox b = BoundingBox::fromPrimitive(*prims[0]);
    for (size_t i = 1; i < prims.size(); ++i) b = b.merge(BoundingBox::fromPrimitive(*prims[i]));
    return b;
}

// ------------------------------------------------------------------ bvh.ts
struct BVHNode {
    BoundingBox bounds;
    PrimList primitives;
    std::unique_ptr<BVHNode> left, right;
    bool isLeaf() const { return !left && !right; }
};
struct BVHInterval {
    double tEnter, tExit;
    const BVHNode* node;
};
struct BVHState {
    std::vector<BVHInterval> intervals;
    size_t currentIntervalIdx = 0;
    bool terminate = false;
};

struct BVH {
    std::unique_ptr<BVHNode> root;
    int maxDepth = 20, maxPrimitivesPerNode = 2;

    explicit BVH(const PrimList& prims) {  // :29-42 (the ctor's `bounds` argument is ignored there too)
        BoundingBox primitiveBounds = computeBounds(prims);
        root = buildNode(prims, primitiveBounds, 0);
    }
    std::unique_ptr<BVHNode> buildNode(const PrimList& prims, const BoundingBox& bounds, int depth) {  // :44-92
        auto node = std::make_unique<BVHNode>();
        node->bounds = bounds;
        if (depth >= maxDepth || (int)prims.size() <= maxPrimitivesPerNode) {
            node->primitives = prims;
            return node;
        }
        vec3 size = glm::v3_create();
        glm::v3_subtract(size, bounds.max, bounds.min);
        int axis = 0;
        if (size[1] > size[0]) axis = 1;
        if (size[2] > size[axis]) axis = 2;
        // [...primitives].sort((a,b) => aPos - bPos): Array.prototype.sort is stable (ES2019).
        std::vector<std::pair<double, const Primitive*>> keyed;
        keyed.reserve(prims.size());
        for (auto* p : prims) keyed.emplace_back(p->getWorldPosition()[axis], p);
        std::stable_sort(keyed.begin(), keyed.end(),
                         [](const auto& a, const auto& b) { return a.first < b.first; });
        size_t mid = keyed.size() / 2;
        PrimList leftStuff, rightStuff;
        for (size_t i = 0; i < mid; ++i) leftStuff.push_back(keyed[i].second);
        for (size_t i = mid; i < keyed.size(); ++i) rightStuff.push_back(keyed[i].second);
        if (leftStuff.empty() || rightStuff.empty()) {
            node->primitives = prims;
            return node;
        }
        BoundingBox leftBounds = computeBounds(leftStuff);
        BoundingBox rightBounds = computeBounds(rightStuff);
        node->left = buildNode(leftStuff, leftBounds, depth + 1);
        node->right = buildNode(rightStuff, rightBounds, depth + 1);
        return node;
    }
    // :95-121  (Set keeps insertion order; a primitive lives in exactly one leaf, so no duplicates arise)
    void queryNode(const BVHNode* node, const vec3& point, PrimList& results) const {
        if (!node->bounds.contains(point)) return;
        if (node->isLeaf()) {
            for (auto* p : node->primitives)
                if (std::find(results.begin(), results.end(), p) == results.end()) results.push_back(p);
            return;
        }
        if (node->left) queryNode(node->left.get(), point, results);
        if (node->right) queryNode(node->right.get(), point, results);
    }
    PrimList getPrimitivesAt(const vec3& point) const {
        PrimList r;
        queryNode(root.get(), point, r);
        return r;
    }
    // :126-178
    std::vector<BVHInterval> findRayIntersections(const vec3& origin, const vec3& direction, double tMin,
                                                  double tMax) const {
        std::vector<BVHInterval> out;
        std::vector<const BVHNode*> stack{root.get()};
        while (!stack.empty()) {
            const BVHNode* node = stack.back();
            stack.pop_back();
            double tEnter, tExit;
            if (!node->bounds.intersectRay(origin, direction, tEnter, tExit)) continue;
            if (tExit < tMin || tEnter > tMax) continue;
            double clampedEnter = js::max2(tEnter, tMin);
            double clampedExit = js::min2(tExit, tMax);
            if (!node->isLeaf()) {
                if (node->left) stack.push_back(node->left.get());
                if (node->right) stack.push_back(node->right.get());
            } else if (!node->primitives.empty()) {
                out.push_back({clampedEnter, clampedExit, node});
            }
        }
        std::stable_sort(out.begin(), out.end(),
                         [](const BVHInterval& a, const BVHInterval& b) { return a.tEnter < b.tEnter; });
        return out;
    }
    // :181-202
    BVHState onRayMarchStart(const vec3& origin, const vec3& dir, double maxDistance) const {
        BVHState s;
        s.intervals = findRayIntersections(origin, dir, 0, maxDistance);
        s.terminate = s.intervals.empty();
        return s;
    }
    // :204-240
    double onRayMarchStep(double currentDistance, BVHState& st) const {
        if (st.terminate) return -1;
        if (st.currentIntervalIdx >= st.intervals.size()) return -1;
        const BVHInterval& cur = st.intervals[st.currentIntervalIdx];
        if (currentDistance < cur.tEnter) return cur.tEnter - currentDistance;
        if (currentDistance > cur.tExit) {
            st.currentIntervalIdx++;
            if (st.currentIntervalIdx < st.intervals.size()) {
                const BVHInterval& nx = st.intervals[st.currentIntervalIdx];
                if (nx.tEnter > currentDistance) return nx.tEnter - currentDistance;
            } else {
                return -1;
            }
        }
        return 0;
    }
};

// ------------------------------------------------------------------ octree.ts
struct OctreeNode {
    BoundingBox bounds;
    PrimList primitives;
    std::vector<std::unique_ptr<OctreeNode>> children;
    bool hasChildren = false;  // children !== null
    int level = 0;
    bool isEmpty = true;
    double minDistance = 0;
    bool isLeaf() const { return !hasChildren; }
};

struct Octree {
    std::unique_ptr<OctreeNode> root;
    int maxDepth = 6, maxPrimitivesPerNode = 4;
    std::vector<BoundingBox> primitiveBounds;

    Octree(const PrimList& prims, const BoundingBox& bounds) {  // :36-50
        for (auto* p : prims) primitiveBounds.push_back(BoundingBox::fromPrimitive(*p));
        root = buildNode(prims, bounds, 0);
        computeMinDistances(root.get());
    }
    std::unique_ptr<OctreeNode> buildNode(const PrimList& prims, const BoundingBox& bounds, int depth) {  // :52-118
        auto node = std::make_unique<OctreeNode>();
        node->bounds = bounds;
        node->level = depth;
        if (depth >= maxDepth || (int)prims.size() <= maxPrimitivesPerNode) {
            node->primitives = prims;
            return node;
        }
        vec3 center = bounds.center();
        std::vector<BoundingBox> childBounds;
        for (int zS = 0; zS < 2; ++zS)
            for (int yS = 0; yS < 2; ++yS)
                for (int xS = 0; xS < 2; ++xS) {
                    double minX = xS == 0 ? bounds.min[0] : center[0];
                    double maxX = xS == 0 ? center[0] : bounds.max[0];
                    double minY = yS == 0 ? bounds.min[1] : center[1];
                    double maxY = yS == 0 ? center[1] : bounds.max[1];
                    double minZ = zS == 0 ? bounds.min[2] : center[2];
                    double maxZ = zS == 0 ? center[2] : bounds.max[2];
                    childBounds.emplace_back(glm::v3_from(minX, minY, minZ), glm::v3_from(maxX, maxY, maxZ));
                }
        std::vector<PrimList> childPrims(8);
        for (auto* p : prims) {
            BoundingBox pb = BoundingBox::fromPrimitive(*p);
            for (int i = 0; i < 8; ++i)
                if (childBounds[i].intersects(pb)) childPrims[i].push_back(p);
        }
        node->hasChildren = true;
        for (int i = 0; i < 8; ++i) {
            if (!childPrims[i].empty()) {
                node->children.push_back(buildNode(childPrims[i], childBounds[i], depth + 1));
            } else {
                auto e = std::make_unique<OctreeNode>();
                e->bounds = childBounds[i];
                e->level = depth + 1;
                node->children.push_back(std::move(e));
            }
        }
        return node;
    }
    double minDistToPrimBounds(const OctreeNode* node) const {
        double minD = js::kInf;
        for (auto& pb : primitiveBounds) {
            double d = node->bounds.distanceToBox(pb);
            if (d < minD) minD = d;
        }
        return minD != js::kInf ? js::max2(0, minD) : 0;
    }
    bool computeMinDistances(OctreeNode* node) {  // :149-191
        if (node->isLeaf()) {
            bool hasPrims = !node->primitives.empty();
            node->isEmpty = !hasPrims;
            node->minDistance = !hasPrims ? minDistToPrimBounds(node) : 0;
            return hasPrims;
        }
        bool subtreeHasPrims = false;
        for (auto& c : node->children)
            if (computeMinDistances(c.get())) subtreeHasPrims = true;
        node->isEmpty = !subtreeHasPrims;
        node->minDistance = node->isEmpty ? minDistToPrimBounds(node) : 0;
        return subtreeHasPrims;
    }
    // :195-220  (tMin/tMax live in Float32Arrays -> f32 rounding; no parallel-ray guard)
    bool intersectRayBox(const vec3& o, const vec3& d, const BoundingBox& box, double& tEnterOut,
                         double& tExitOut) const {
        vec3 tMin = glm::v3_create(), tMax = glm::v3_create();
        for (int i = 0; i < 3; ++i) {
            double invD = 1.0 / d[i];
            double t0 = (box.min[i] - o[i]) * invD;
            double t1 = (box.max[i] - o[i]) * invD;
            if (invD < 0.0) std::swap(t0, t1);
            tMin.e[i] = js::f32(t0);
            tMax.e[i] = js::f32(t1);
        }
        double tEnter = js::max3(tMin[0], tMin[1], tMin[2]);
        double tExit = js::min3(tMax[0], tMax[1], tMax[2]);
        if (tEnter > tExit || tExit < 0) return false;
        tEnterOut = js::max2(0, tEnter);
        tExitOut = tExit;
        return true;
    }
    // :223-248
    const OctreeNode* findNodeRecursive(const OctreeNode* node, const vec3& p) const {
        if (!node->bounds.contains(p)) return nullptr;
        if (node->isLeaf() || node->level == maxDepth) return node;
        for (auto& c : node->children) {
            const OctreeNode* f = findNodeRecursive(c.get(), p);
            if (f) return f;
        }
        return node;
    }
    const OctreeNode* findNode(const vec3& p) const { return findNodeRecursive(root.get(), p); }
    // :252-278
    double marchRay(const vec3& o, const vec3& d, double currentDist) const {
        vec3 currentPos = glm::v3_create();
        glm::v3_scale_and_add(currentPos, o, d, currentDist);
        const OctreeNode* node = findNode(currentPos);
        if (!node) return 0;
        if (node->isEmpty) {
            double tEnter, tExit;
            if (intersectRayBox(o, d, node->bounds, tEnter, tExit)) {
                double toExit = js::max2(0, tExit - currentDist);
                double step = js::max2(0, js::min2(toExit, node->minDistance * 0.99));
                return step > 0 ? step + 0.001 : 0;
            }
        }
        return 0;
    }
};

// ------------------------------------------------------------------ camera.ts
struct Camera {
    mat4 orbitCentre = glm::m4_create();
    double cameraDistance = 3;
    double pitch = 0, yaw = 0;
    mat4 cameraTransform = glm::m4_create();
    Camera() { updateCameraTransform(); }
    void setAngles(double p, double y) {  // :58-62
        pitch = js::min2(js::max2(p, -M_PI / 2), M_PI / 2);
        yaw = y;
        updateCameraTransform();
    }
    void rotateCamera(double p, double y) {  // :26-31
        pitch = js::min2(js::max2(pitch + p, -M_PI / 2), M_PI / 2);
        yaw += y;
        updateCameraTransform();
    }
    void getRotationMatrix(mat4& out) const {  // :38-44
        out = glm::m4_create();
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) out.e[4 * c + r] = cameraTransform.e[4 * c + r];
    }
    void getPosition(vec3& out) const {  // :64-69
        out.e[0] = cameraTransform.e[12];
        out.e[1] = cameraTransform.e[13];
        out.e[2] = cameraTransform.e[14];
    }
    void updateCameraTransform() {  // :81-88
        mat4 tempOrbitCentre = glm::m4_create();
        glm::m4_rotate_y(tempOrbitCentre, glm::m4_create(), yaw);
        glm::m4_rotate_x(orbitCentre, tempOrbitCentre, pitch);
        glm::m4_translate(cameraTransform, orbitCentre, glm::v3_from(0, 0, std::fabs(cameraDistance)));
    }
};

// ------------------------------------------------------------------ sceneManager.ts
inline mat4 getTransform(double x, double y, double z, const vec3* rotation) {  // :21-37
    mat4 model = glm::m4_create();
    if (rotation) {
        glm::m4_from_translation(model, x, y, z);
        glm::m4_rotate_x(model, model, (*rotation)[0]);
        glm::m4_rotate_y(model, model, (*rotation)[1]);
        glm::m4_rotate_z(model, model, (*rotation)[2]);
    } else {
        const double q[4] = {0, 0, 0, 1}, v[3] = {x, y, z}, s[3] = {1, 1, 1};
        glm::m4_from_rts(model, q, v, s);
    }
    mat4 worldToLocal = glm::m4_create();
    glm::m4_invert(worldToLocal, model);
    return worldToLocal;
}
inline Primitive createSphere(double x, double y, double z, double radius, const vec3* rot = nullptr) {  // :39-41
    Primitive p;
    p.type = SPHERE;
    p.transform = getTransform(x, y, z, rot);
    p.radius = radius;
    return p;
}
inline Primitive createBox(double x, double y, double z, const vec3& half, const vec3* rot = nullptr) {  // :43-45
    Primitive p;
    p.type = BOX;
    p.transform = getTransform(x, y, z, rot);
    p.halfSize = half;
    return p;
}
inline Primitive createTorus(double x, double y, double z, double radius, const vec3* rot = nullptr) {  // :47-49
    Primitive p;
    p.type = TORUS;
    p.transform = getTransform(x, y, z, rot);
    p.majorRadius = radius;
    p.minorRadius = radius / 4;
    return p;
}

// sceneManager.ts:73-100 : operator factories
inline std::shared_ptr<Primitive> boxed(const Primitive& p) { return std::make_shared<Primitive>(p); }
inline Primitive createSmoothUnion(const Primitive& prim1, const Primitive& prim2, double k) {  // smoothUnion.ts:9-15
    Primitive p;
    p.type = SMOOTH_UNION;
    p.transform = glm::m4_create();
    p.a = boxed(prim1);
    p.b = boxed(prim2);
    p.smoothness = k;
    return p;
}
inline Primitive createSmoothSubtract(const Primitive& prim1, const Primitive& prim2, double k) {  // smoothSubstraction.ts:9-14
    Primitive p = createSmoothUnion(prim1, prim2, k);
    p.type = SMOOTH_SUBTRACTION;
    return p;
}
inline Primitive createTwist(const Primitive& prim, double twistAmount) {  // twist.ts:8-12
    Primitive p;
    p.type = TWIST;
    p.transform = prim.transform;
    p.a = boxed(prim);
    p.twistAmount = twistAmount;
    return p;
}
inline Primitive createRound(const Primitive& prim, double radius) {  // round.ts:8-12
    Primitive p;
    p.type = ROUND;
    p.transform = prim.transform;
    p.a = boxed(prim);
    p.radius = radius;
    return p;
}
inline Primitive createRepetition(const Primitive& prim, const vec3& spacing) {  // repetition.ts:8-12
    Primitive p;
    p.type = REPETITION;
    p.transform = prim.transform;
    p.a = boxed(prim);
    p.spacing = spacing;
    return p;
}
inline Primitive createAnimatedTranslate(const Primitive& prim, const vec3& direction, double amplitude, double speed) {
    Primitive p;  // animatedTranslate.ts:15-28
    p.type = ANIMATED_TRANSLATE;
    p.transform = prim.transform;
    p.a = boxed(prim);
    glm::v3_normalize(p.direction, direction);
    p.amplitude = amplitude;
    p.speed = speed;
    p.time = 0;
    return p;
}

// sceneManager.ts:51-71
inline Primitive createMandelbulb(double x, double y, double z, double power, int iterations, bool enableAnimation, double animationSpeed,
                                  const vec3* rot = nullptr) {
    Primitive p;
    p.type = MANDELBULB;
    p.transform = getTransform(x, y, z, rot);
    const double half[3] = {0.5, 0.5, 0.5};
    glm::m4_scale(p.transform, p.transform, half);
    p.power = power;
    p.iterations = iterations;
    p.enableAnimation = enableAnimation;
    p.animationSpeed = animationSpeed;
    return p;
}

// all 19 presets (sceneManager.ts:102-356)
inline bool makePreset(int idx, std::vector<Primitive>& out) {
    out.clear();
    switch (idx) {
        case 0: out.push_back(createSphere(0, 0, 0, 1.5)); break;  // :104-109
        case 1:                                                    // :110-121
            out.push_back(createSphere(0.8, -0.3, 0.2, 0.4));
            out.push_back(createSphere(-0.5, 0.9, -0.1, 0.5));
            out.push_back(createSphere(0.2, 0.1, 0.8, 0.3));
            out.push_back(createSphere(-0.9, -0.4, -0.6, 0.6));
            out.push_back(createSphere(0.4, -0.8, 0.5, 0.35));
            out.push_back(createSphere(-0.2, 0.6, -0.9, 0.4));
            out.push_back(createSphere(0.7, 0.3, -0.4, 0.25));
            break;
        case 2:  // :122-135
            for (int y = -1; y <= 1; ++y)
                for (int x = -1; x <= 1; ++x) out.push_back(createSphere(x, y, 0, 0.3));
            break;
        case 3: {  // :136-158
            const int gridSize = 5;
            const double spacing = 0.6;
            const double offset = (gridSize - 1) * spacing / 2;
            for (int x = 0; x < gridSize; ++x)
                for (int y = 0; y < gridSize; ++y)
                    for (int z = 0; z < gridSize; ++z)
                        out.push_back(createSphere(x * spacing - offset, y * spacing - offset, z * spacing - offset, 0.15));
            break;
        }
        case 4:  // :159-170
            out.push_back(createSphere(0, 0, 0, 0.5));
            out.push_back(createSphere(1.2, 0, 0, 0.3));
            out.push_back(createSphere(-1.2, 0, 0, 0.3));
            out.push_back(createSphere(0, 1.2, 0, 0.3));
            out.push_back(createSphere(0, -1.2, 0, 0.3));
            out.push_back(createSphere(0, 0, 1.2, 0.3));
            out.push_back(createSphere(0, 0, -1.2, 0.3));
            break;
        case 5: {  // :171-176
            vec3 rot = glm::v3_from(-M_PI / 2, 0, 0);
            out.push_back(createTorus(0, 0, 0, 1.3, &rot));
            break;
        }
        case 7: out.push_back(createBox(0, 0, 0, glm::v3_from(1, 1, 1))); break;  // :187-192
        case 8:                                                                  // :193-199
            out.push_back(createSphere(-0.7, 0, 0, 0.5));
            out.push_back(createBox(1, 0, 0, glm::v3_from(0.5, 0.5, 0.5)));
            break;
        case 9:  // :200-207
            out.push_back(createBox(0, 0.5, 0, glm::v3_from(0.9, 0.25, 0.9)));
            out.push_back(createBox(0, 0, 0, glm::v3_from(0.6, 0.25, 0.6)));
            out.push_back(createBox(0, -0.5, 0, glm::v3_from(0.3, 0.25, 0.3)));
            break;
        case 6:  // :178-186 Rounded Box
            out.push_back(createRound(createBox(0, 0, 0, glm::v3_from(0.4, 0.4, 0.4)), 0.3));
            break;
        case 10:  // :209-218 Smooth Union
            out.push_back(createSmoothUnion(createSphere(0, 0, 0, 0.5), createBox(0, 0.5, 0, glm::v3_from(1, 0.2, 1)), 0.2));
            break;
        case 11: {  // :219-231 Smooth Subtraction
            vec3 rot = glm::v3_from(0, M_PI / 4, 0);
            out.push_back(createSmoothSubtract(createRound(createBox(0, 0, 0, glm::v3_from(1, 1, 1), &rot), 0.1),
                                               createSphere(0, 0, 0, 0.9), 0.2));
            break;
        }
        case 12:  // :233-246 Smooth Union [A]
            out.push_back(createSmoothUnion(createAnimatedTranslate(createSphere(0, 0, 0, 1), glm::v3_from(1, 0, 0), 3.0, 0.005),
                                            createSphere(0, 0, 0, 1), 0.2));
            break;
        case 13: out.push_back(createMandelbulb(0, 0, 0, 8, 80, true, -0.0001)); break;  // :247-253 Mandelbulb [A]
        case 14: {  // :254-262 Twisted Torus
            vec3 rot = glm::v3_from(-M_PI / 2, 0, 0);
            out.push_back(createTwist(createTorus(0, 0, 0, 1.3, &rot), 3));
            break;
        }
        case 15:  // :263-271 Infinite Spheres
            out.push_back(createRepetition(createSphere(0, 0, 0, 0.3), glm::v3_from(1.5, 1.5, 1.5)));
            break;
        case 16:  // :272-283 Screw
            out.push_back(createRound(createTwist(createBox(0, 0, 0, glm::v3_from(0.4, 1.5, 0.4)), 4.0), 0.1));
            break;
        case 17: {  // :284-325 Chicken: a left-deep chain of nine smooth unions over ten boxes
            struct B { double x, y, z, hx, hy, hz; };
            static const B boxes[10] = {{0, 0, 0, 0.6, 0.6, 0.8},      {0, -0.2, 0, 0.8, 0.4, 0.6},   {0, -0.8, 0.8, 0.4, 0.6, 0.3},
                                        {0, -0.8, 1.2, 0.4, 0.2, 0.2}, {0, -0.4, 1.0, 0.2, 0.2, 0.2}, {0.3, 1, 0, 0.1, 0.6, 0.01},
                                        {-0.3, 1, 0, 0.1, 0.6, 0.01},  {0, 1.6, 0.2, 0.6, 0.01, 0.2}, {0.3, 1.6, 0.5, 0.1, 0.01, 0.1},
                                        {-0.3, 1.6, 0.5, 0.1, 0.01, 0.1}};
            Primitive acc = createBox(boxes[0].x, boxes[0].y, boxes[0].z, glm::v3_from(boxes[0].hx, boxes[0].hy, boxes[0].hz));
            for (int i = 1; i < 10; ++i)
                acc = createSmoothUnion(acc, createBox(boxes[i].x, boxes[i].y, boxes[i].z, glm::v3_from(boxes[i].hx, boxes[i].hy, boxes[i].hz)),
                                        0.0001);
            out.push_back(acc);
            break;
        }
        case 18: {  // :326-356 "67"
            vec3 r5 = glm::v3_from(0, 0, M_PI / 5), r7 = glm::v3_from(0, 0, M_PI / 7), r2 = glm::v3_from(0, 0, M_PI / 2);
            vec3 rt = glm::v3_from(-M_PI / 2, 0, 0);
            out.push_back(createSmoothUnion(createRound(createBox(-1.25, -0.8, 0, glm::v3_from(0.05, 0.7, 0.05), &r5), 0.20),
                                            createRound(createTorus(-1.25, 0.5, 0, 0.8, &rt), 0.05), 0.0001));
            out.push_back(createSmoothUnion(createRound(createBox(1.35, 0, 0, glm::v3_from(0.05, 1.5, 0.05), &r7), 0.20),
                                            createRound(createBox(1.25, -1.4, 0, glm::v3_from(0.05, 0.8, 0.05), &r2), 0.20), 0.0001));
            break;
        }
        default: return false;
    }
    for (size_t i = 0; i < out.size(); ++i) out[i].index = (int)i;
    return true;
}

// mulberry32: pure 32-bit integer ops.
struct Mulberry32 {
    uint32_t a;
    explicit Mulberry32(uint32_t seed) : a(seed) {}
    double next() {
        a += 0x6D2B79F5u;
        uint32_t t = (a ^ (a >> 15)) * (1u | a);
        t = (t + ((t ^ (t >> 7)) * (61u | t))) ^ t;
        return (double)(t ^ (t >> 14)) / 4294967296.0;
    }
};
// config 4: prims 0-6 = preset 1, the rest seeded spheres.
inline void makeSyntheticSpheres(int n, uint32_t seed, std::vector<Primitive>& out) {
    makePreset(1, out);
    if ((int)out.size() > n) out.resize(n);
    Mulberry32 rng(seed);
    while ((int)out.size() < n) {
        double cx = -2.5 + 5 * rng.next();
        double cy = -2.5 + 5 * rng.next();
        double cz = -2.5 + 5 * rng.next();
        double r = 0.02 + 0.03 * rng.next();
        out.push_back(createSphere(cx, cy, cz, r));
    }
    for (size_t i = 0; i < out.size(); ++i) out[i].index = (int)i;
}

// ------------------------------------------------------------------ scene.ts
enum AccelKind { ACCEL_NONE = 0, ACCEL_OCTREE = 1, ACCEL_BVH = 2 };

struct Scene {
    std::vector<Primitive> objects;  // objectSDFs
    PrimList objectSDFs;
    Camera camera;
    int accel = ACCEL_NONE;
    std::unique_ptr<Octree> octree;
    std::unique_ptr<BVH> bvh;

    void setObjects(std::vector<Primitive> prims) {
        objects = std::move(prims);
        objectSDFs.clear();
        for (size_t i = 0; i < objects.size(); ++i) {
            objects[i].index = (int)i;
            objectSDFs.push_back(&objects[i]);
        }
        octree.reset();
        bvh.reset();
    }
    // scene.ts:48-85
    void buildAccel(int kind) {
        accel = kind;
        octree.reset();
        bvh.reset();
        if (kind == ACCEL_OCTREE) {
            BoundingBox bounds(glm::v3_from(-10, -10, -10), glm::v3_from(10, 10, 10));
            octree = std::make_unique<Octree>(objectSDFs, bounds);
        } else if (kind == ACCEL_BVH) {
            bvh = std::make_unique<BVH>(objectSDFs);
        }
    }
    // scene.ts:135-140 (called once per job by raymarcher.ts:59)
    void updateTime(double time) {
        for (auto& p : objects) p.setTime(time);
    }
    // scene.ts:144-190
    double getDistance(const vec3& position, uint32_t& count) const {
        const double MAX_DIST = 10;
        double closest = MAX_DIST;
        if (accel == ACCEL_OCTREE && octree) {
            const OctreeNode* node = octree->findNode(position);
            if (node) {
                if (!node->primitives.empty()) {
                    for (auto* p : node->primitives) {
                        count++;
                        closest = js::min2(p->sdf(position), closest);
                    }
                } else if (node->isEmpty) {
                    const double safety = 0.99;
                    closest = js::min2(closest, node->minDistance * safety);
                }
                return closest;
            }
        } else if (accel == ACCEL_BVH && bvh) {
            PrimList cands = bvh->getPrimitivesAt(position);
            const PrimList& use = cands.empty() ? objectSDFs : cands;
            for (auto* p : use) {
                count++;
                closest = js::min2(p->sdf(position), closest);
            }
            return closest;
        }
        for (auto* p : objectSDFs) {
            count++;
            closest = js::min2(p->sdf(position), closest);
        }
        return closest;
    }
};

}  // namespace orc
