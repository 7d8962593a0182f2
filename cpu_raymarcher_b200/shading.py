"""ShadingModel mirrors (src/util/shading_models/*.ts): shade(depth, normal, sdfEval, iters, w, h) -> RGBA.
Each call runs the shade kernel of librm_b200.so (rm_shade); the fused variant is rm_request.shader."""
from __future__ import annotations

from .renderer import Context


class ShadingModel:
    name = "normal"

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def shade(self, depth, normal, sdf_eval, iters, width, height):  # shadingModel.ts:8-16
        return self.ctx.shade(self.name, depth, normal, sdf_eval, iters, width, height)


class NormalModel(ShadingModel):
    name = "normal"


class PhongModel(ShadingModel):
    name = "phong"


class SDFHeatmap(ShadingModel):
    name = "sdf-heatmap"


class IterationHeatmap(ShadingModel):
    name = "iteration-heatmap"


def create_shading_model_from_value(ctx: Context, value: str) -> ShadingModel:  # main.ts:33-45
    return {"phong": PhongModel, "sdf-heatmap": SDFHeatmap, "iteration-heatmap": IterationHeatmap}.get(value, NormalModel)(ctx)
