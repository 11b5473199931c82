"""Test-side helper: drives the product exactly as main.cpp drives the reference (ComputeMaterial / ComputeModel)."""
import numpy as np

import vulkan_compute_ray_tracing_b200 as vcrt


class GpuScene:
    """initScene() of main.cpp:76-153, headless: 1 UBO bundle, 2 storage images, 5 storage buffer bundles."""

    def __init__(self, scene, w, h, shader="ray-trace-compute", device=0):
        self.scene, self.w, self.h = scene, w, h
        self.ubo = vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), bytes(32))
        self.target, self.accum = vcrt.Image(w, h), vcrt.Image(w, h)
        m = vcrt.ComputeMaterial("resources/shaders/generated/%s.spv" % shader, device=device)
        m.addUniformBufferBundle(self.ubo, vcrt.VK_SHADER_STAGE_COMPUTE_BIT)
        m.addStorageImage(self.target, vcrt.VK_SHADER_STAGE_COMPUTE_BIT)
        m.addStorageImage(self.accum, vcrt.VK_SHADER_STAGE_COMPUTE_BIT)
        for name in ("triangles", "materials", "bvh", "lights", "spheres"):
            m.addStorageBufferBundle(vcrt.BufferUtils.createBundle(vcrt.BufferBundle(1), scene[name]), vcrt.VK_SHADER_STAGE_COMPUTE_BIT)
        self.material = m
        self.model = vcrt.ComputeModel(m)

    def set_camera(self, cam, sample=0, num_triangles=None):
        """updateScene() of main.cpp:166-183: write the 32-byte UBO into the bundle's buffer."""
        self.model.getMaterial().getUniformBufferBundles()[0].data.buffers[0].write(vcrt.pack_ubo(cam, sample, self.scene, num_triangles=num_triangles))

    def frames(self, cam, n, full_cover=True):
        """The reference frame loop: one computeCommand per frame with currentSample = frame index."""
        gx = (self.w + 31) // 32 if full_cover else self.w // 32
        gy = (self.h + 31) // 32 if full_cover else self.h // 32
        for s in range(n):
            self.set_camera(cam, s)
            self.model.computeCommand(None, 0, gx, gy, 1)
        return self.target.read()

    def render(self, cam, **kw):
        want_aov = kw.pop("want_aov", False)
        clear = kw.pop("clear", True)
        num_triangles = kw.pop("num_triangles", None)
        kw["flags"] = kw.get("flags", 0) | (vcrt.FLAG_WRITE_AOV if want_aov else 0)
        p = vcrt.render_params(**kw)
        self.set_camera(cam, 0, num_triangles)
        if clear:
            self.material.clearAccum()
        self.material.resetCounters()
        self.model.renderCommand(None, 0, p)
        out = dict(counters=self.material.counters())
        if p.accum_mode == vcrt.ACCUM["f32"]:
            out["accumf"] = self.material.readAccumF32()
        else:
            out["target"] = self.target.read()
            out["accum8"] = self.accum.read()
        if want_aov:
            out["aov"] = self.material.readAov()
        return out

    def close(self):
        self.material.destroy()
