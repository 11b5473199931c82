"""Host-side mirror of the reference's interface for the hot path (Python over the C ABI).

Same names, argument meaning and error behaviour as the reference's C++ classes:
  BufferBundle / BufferUtils.createBundle   src/memory/Buffer.h:42-108
  Image                                     src/memory/Image.h (storage images, main.cpp:108-140)
  ComputeMaterial                           src/scene/ComputeMaterial.{h,cpp} (+ the add*/get* surface of Material.h:29-43)
  ComputeModel                              src/scene/ComputeModel.{h,cpp}
The Vulkan bodies (descriptor sets, pipeline, vkCmdDispatch) are replaced by calls into libvcrt.so.
Errors surface as VcrtError where the reference throws std::runtime_error.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import Counters, RenderParams, Ubo, VcrtError

VK_SHADER_STAGE_COMPUTE_BIT = 0x20

SHADER = {"full": 0, "simple": 1}
TRAVERSAL = {"reference": 0, "fast": 1, "brute_force": 2}
RNG = {"pcg_ref": 0, "philox": 1}
ACCUM = {"rgba8_ref": 0, "f32": 1}
TRIG = {"libm": 0, "portable": 1}
FLAG_REF_DISPATCH_COVERAGE, FLAG_WRITE_AOV, FLAG_COUNT_TRAVERSAL, FLAG_STATIC_KERNEL, FLAG_MEGAKERNEL, FLAG_WAVEFRONT = 1, 2, 4, 8, 16, 32

AOV_DTYPE = np.dtype([("triangle", "<i4"), ("material", "<i4"), ("t", "<f4"), ("backFace", "<u4")])
RECORD_BYTES = {"triangles": 48, "materials": 32, "bvh": 48, "lights": 8, "spheres": 32}


def render_params(shader=None, traversal="fast", rng="pcg_ref", accum="f32", trig="libm", max_bounces=0, stack_depth=0,
                  lights_length=0, sample_begin=0, sample_count=1, tile_rank=0, tile_count=0, philox_seed=0, flags=0):
    """vcrt_render_params with the shader's compile-time constants as fields (0 = the shader's own value)."""
    p = RenderParams()
    p.struct_size = C.sizeof(RenderParams)
    p.shader = 0 if shader is None else (SHADER[shader] if isinstance(shader, str) else int(shader))
    p.traversal = TRAVERSAL[traversal] if isinstance(traversal, str) else int(traversal)
    p.rng_mode = RNG[rng] if isinstance(rng, str) else int(rng)
    p.accum_mode = ACCUM[accum] if isinstance(accum, str) else int(accum)
    p.trig_mode = TRIG[trig] if isinstance(trig, str) else int(trig)
    p.max_bounces, p.stack_depth, p.lights_length = max_bounces, stack_depth, lights_length
    p.sample_begin, p.sample_count, p.tile_rank, p.tile_count = sample_begin, sample_count, tile_rank, tile_count
    p.philox_seed, p.flags = philox_seed, flags
    return p


class Buffer:
    """One host-visible buffer (Buffer.h:15-40): `data` is the mapped memory the host writes."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8)).copy() if not isinstance(data, np.ndarray) \
            else np.ascontiguousarray(data).view(np.uint8).reshape(-1).copy()
        self.size = self.data.nbytes

    def write(self, blob):
        """vmaMapMemory + memcpy + vmaUnmapMemory (main.cpp:176-180)."""
        b = np.frombuffer(bytes(blob), dtype=np.uint8)
        if b.nbytes != self.size:
            raise VcrtError("failed to write buffer: size mismatch")
        self.data[:] = b


class BufferBundle:
    """N identical buffers, one per swapchain image (Buffer.h:42-53); headless default N = 1."""

    def __init__(self, bundleSize=1):
        self.buffers = [None] * bundleSize


class BufferUtils:
    @staticmethod
    def createBundle(bufferBundle, data, usage=None, memoryUsage=None):
        """BufferUtils::createBundle<T> (Buffer.h:95-108): every buffer of the bundle gets a copy of `data`."""
        for i in range(len(bufferBundle.buffers)):
            bufferBundle.buffers[i] = Buffer(data)
        return bufferBundle


class Image:
    """rgba8 storage image (ImageUtils::createImage, main.cpp:108-140).  Filled in by ComputeMaterial.init()."""

    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)
        self._material = None
        self._slot = None

    def read(self):
        """Device -> host copy of the texels, (H, W, 4) uint8."""
        if self._material is None:
            raise VcrtError("failed to read image: not bound to an initialised ComputeMaterial")
        return self._material._read_image(self._slot)


class PinnedFrame:
    """An (H, W, 4) uint8 frame in page-locked host memory (vcrt_alloc_host): the destination of a read-back that overlaps the
    rendering of later frames (ComputeMaterial.frameSubmit)."""

    def __init__(self, width, height):
        self._ptr = C.c_void_p()
        nbytes = int(width) * int(height) * 4
        if N.lib().vcrt_alloc_host(nbytes, C.byref(self._ptr)) != 0:
            raise VcrtError(N.lib().vcrt_last_error(None).decode())
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_uint8)), shape=(int(height), int(width), 4))

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            N.lib().vcrt_free_host(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Descriptor:
    """Descriptor<T> (Material.h:12-17)."""

    def __init__(self, data, shaderStageFlags):
        self.data, self.shaderStageFlags = data, shaderStageFlags


class ComputeMaterial:
    """mcvkp::ComputeMaterial (ComputeMaterial.h:12-26).

    Binding numbers follow the reference: uniform buffers first, then storage images, then storage buffers, each
    class in insertion order (Material.cpp:258-311) -- i.e. 0 UBO, 1 target, 2 accumulation, 3 triangles,
    4 materials, 5 bvh, 6 lights, 7 spheres for the call sequence of main.cpp:145-152.
    """

    def __init__(self, computeShaderPath, device=0):
        self.m_computeShaderPath = computeShaderPath
        self.m_device = device
        self.m_uniformBufferBundleDescriptors = []
        self.m_storageBufferBundleDescriptors = []
        self.m_storageImageDescriptors = []
        self.m_initialized = False
        self._ctx = None

    # ---- Material.h:29-43
    def addUniformBufferBundle(self, bufferBundle, shaderStageFlags=VK_SHADER_STAGE_COMPUTE_BIT):
        self.m_uniformBufferBundleDescriptors.append(Descriptor(bufferBundle, shaderStageFlags))

    def addStorageImage(self, image, shaderStageFlags=VK_SHADER_STAGE_COMPUTE_BIT):
        self.m_storageImageDescriptors.append(Descriptor(image, shaderStageFlags))

    def addStorageBufferBundle(self, bufferBundle, shaderStageFlags=VK_SHADER_STAGE_COMPUTE_BIT):
        self.m_storageBufferBundleDescriptors.append(Descriptor(bufferBundle, shaderStageFlags))

    def getUniformBufferBundles(self):
        return self.m_uniformBufferBundleDescriptors

    def getStorageBufferBundles(self):
        return self.m_storageBufferBundleDescriptors

    def getStorageImages(self):
        return self.m_storageImageDescriptors

    # ---- ComputeMaterial.cpp:15-26
    def init(self):
        if self.m_initialized:
            return
        L = N.lib()
        if len(self.m_uniformBufferBundleDescriptors) != 1 or len(self.m_storageImageDescriptors) != 2 or \
                len(self.m_storageBufferBundleDescriptors) != 5:
            raise VcrtError("failed to create compute pipeline layout: the kernel expects 1 uniform buffer, 2 storage images "
                            "and 5 storage buffers (ray-trace-compute.comp:8-39)")
        ctx = C.c_void_p()
        if L.vcrt_create(self.m_device, C.byref(ctx)) != 0:
            raise VcrtError(L.vcrt_last_error(None).decode())
        self._ctx = ctx
        self._check(L.vcrt_set_shader(ctx, self.m_computeShaderPath.encode()))
        target, accum = (d.data for d in self.m_storageImageDescriptors)
        if (target.width, target.height) != (accum.width, accum.height):
            raise VcrtError("failed to create descriptor sets: target and accumulation images differ in size")
        self._check(L.vcrt_set_image_size(ctx, target.width, target.height))
        target._material, target._slot = self, 0
        accum._material, accum._slot = self, 1
        for i, d in enumerate(self.m_storageBufferBundleDescriptors):
            buf = d.data.buffers[0]
            self._check(L.vcrt_set_buffer(ctx, 3 + i, buf.data.ctypes.data if buf.size else None, buf.size))
        self.m_initialized = True

    def bind(self, commandBuffer, currentFrame):
        """vkCmdBindPipeline + vkCmdBindDescriptorSets (ComputeMaterial.cpp:63-68): the descriptor set of frame
        `currentFrame` points at buffers[currentFrame] of the uniform bundle -- its 32 bytes become the kernel's UBO."""
        self._require()
        bundle = self.m_uniformBufferBundleDescriptors[0].data
        buf = bundle.buffers[currentFrame]
        if buf.size != 32:
            raise VcrtError("failed to bind uniform buffer: expected 32 bytes")
        ubo = Ubo.from_buffer_copy(buf.data.tobytes())
        self._check(N.lib().vcrt_set_ubo(self._ctx, C.byref(ubo)))

    # ---- extensions used by the headless driver / tests / bench
    def _require(self):
        if not self.m_initialized:
            raise VcrtError("failed to use compute material: init() has not run")

    def _check(self, rc):
        if rc != 0:
            raise VcrtError(N.lib().vcrt_last_error(self._ctx).decode())

    def _read_image(self, slot):
        self._require()
        t = self.m_storageImageDescriptors[slot].data
        out = np.empty((t.height, t.width, 4), np.uint8)
        fn = N.lib().vcrt_read_target_rgba8 if slot == 0 else N.lib().vcrt_read_accum_rgba8
        self._check(fn(self._ctx, out.ctypes.data, out.nbytes))
        return out

    def updateStorageBuffer(self, index, data):
        """Re-upload storage buffer `index` (0 = triangles ... 4 = spheres) from host memory."""
        self._require()
        buf = Buffer(data)
        self.m_storageBufferBundleDescriptors[index].data.buffers[0] = buf
        self._check(N.lib().vcrt_set_buffer(self._ctx, 3 + index, buf.data.ctypes.data if buf.size else None, buf.size))

    def readAccumF32(self):
        self._require()
        t = self.m_storageImageDescriptors[0].data
        out = np.empty((t.height, t.width, 4), np.float32)
        self._check(N.lib().vcrt_read_accum_f32(self._ctx, out.ctypes.data, out.nbytes))
        return out

    def writeAccumF32(self, arr):
        self._require()
        arr = np.ascontiguousarray(arr, np.float32)
        self._check(N.lib().vcrt_write_accum_f32(self._ctx, arr.ctypes.data, arr.nbytes))

    def readAov(self):
        self._require()
        t = self.m_storageImageDescriptors[0].data
        out = np.empty((t.height, t.width), AOV_DTYPE)
        self._check(N.lib().vcrt_read_aov(self._ctx, out.ctypes.data, out.nbytes))
        return out

    def clearAccum(self):
        self._require()
        self._check(N.lib().vcrt_clear_accum(self._ctx))

    def resolve(self, total_samples, gamma=0.0):
        self._require()
        self._check(N.lib().vcrt_resolve(self._ctx, int(total_samples), float(gamma)))

    def postProcess(self, mix=0.0, sigma=2.0, k_sigma=2.0, threshold=0.05, gamma=2.2):
        """post-process-shader.frag on the target image -> (H, W, 4) uint8 present image.  Defaults = the shipped shader
        (denoiser off, gamma 2.2); mix=0.5 enables smartDeNoise as its commented-out call would (:64)."""
        self._require()
        self._check(N.lib().vcrt_post_process(self._ctx, mix, sigma, k_sigma, threshold, gamma))
        t = self.m_storageImageDescriptors[0].data
        out = np.empty((t.height, t.width, 4), np.uint8)
        self._check(N.lib().vcrt_read_present_rgba8(self._ctx, out.ctypes.data, out.nbytes))
        return out

    def devicePtr(self, what):
        """(pointer, bytes) of a ctx-owned image: 0 target rgba8, 1 accumulation rgba8, 2 f32 accumulation, 3 AOV."""
        self._require()
        p, n = C.c_void_p(), C.c_size_t()
        self._check(N.lib().vcrt_device_ptr(self._ctx, what, C.byref(p), C.byref(n)))
        return p.value, n.value

    def setOption(self, key, value):
        """Tunables that do not change results, e.g. ("fast_bvh", "sah" | "topology")."""
        self._require()
        self._check(N.lib().vcrt_set_option(self._ctx, key.encode(), value.encode()))

    def packTiles(self, what, tile_rank, tile_count, packed_ptr, nbytes):
        """Owned tiles of image `what` (0 target rgba8, 2 f32 accumulation) -> packed device buffer (vcrt_pack_tiles)."""
        self._require()
        self._check(N.lib().vcrt_pack_tiles(self._ctx, what, tile_rank, tile_count, C.c_void_p(packed_ptr), nbytes))

    def unpackTiles(self, what, tile_rank, tile_count, packed_ptr, nbytes):
        self._require()
        self._check(N.lib().vcrt_unpack_tiles(self._ctx, what, tile_rank, tile_count, C.c_void_p(packed_ptr), nbytes))

    def getInfo(self, key):
        """Read-only facts as text, e.g. getInfo("fast_nodes") -> "q15x4" | "q15" | "f32" | "none"."""
        self._require()
        buf = C.create_string_buffer(256)
        self._check(N.lib().vcrt_get_info(self._ctx, key.encode(), buf, 256))
        return buf.value.decode()

    def setStream(self, cuda_stream):
        """Run on a caller-owned CUDA stream (integer cudaStream_t handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._require()
        self._check(N.lib().vcrt_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._require()
        self._check(N.lib().vcrt_synchronize(self._ctx))

    def counters(self):
        self._require()
        c = Counters()
        self._check(N.lib().vcrt_get_counters(self._ctx, C.byref(c)))
        return c

    def resetCounters(self):
        self._require()
        self._check(N.lib().vcrt_reset_counters(self._ctx))

    # ---- frames in flight (main.cpp:68 MAX_FRAMES_IN_FLIGHT, :298-316 fences, :325 vkWaitForFences, :394): vcrt_frames_* / vcrt_frame_*
    def framesBegin(self, frames_in_flight=2):
        self._require()
        self._check(N.lib().vcrt_frames_begin(self._ctx, int(frames_in_flight)))

    def frameSubmit(self, params, total_samples=0, gamma=0.0, out=None, currentFrame=0):
        """One 1-spp frame on the next slot: render, fold into the accumulation in frame order, resolve, read back into `out`
        (a PinnedFrame or a (H, W, 4) uint8 array; None = no read-back).  Returns the slot; `out` is valid after frameWait(slot) and
        must be kept alive by the caller until then (the copy into it is asynchronous).
        The UBO is buffers[currentFrame] of the uniform bundle, as in bind()."""
        self._require()
        if params.shader == 0 and self.m_computeShaderPath.split("/")[-1].split(".")[0] == "ray-trace-compute-simple":
            params.shader = 1
        self.bind(None, currentFrame)
        arr = out.array if isinstance(out, PinnedFrame) else out
        slot = C.c_uint32()
        self._check(N.lib().vcrt_frame_submit(self._ctx, C.byref(params), int(total_samples), float(gamma), arr.ctypes.data if arr is not None else None,
                                              arr.nbytes if arr is not None else 0, C.byref(slot)))
        return slot.value

    def frameWait(self, slot):
        self._require()
        self._check(N.lib().vcrt_frame_wait(self._ctx, int(slot)))

    def framesEnd(self):
        self._require()
        self._check(N.lib().vcrt_frames_end(self._ctx))

    def destroy(self):
        """~Material (Material.cpp:21-29)."""
        if self._ctx is not None:
            N.lib().vcrt_destroy(self._ctx)
            self._ctx = None
            self.m_initialized = False

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class ComputeModel:
    """mcvkp::ComputeModel (ComputeModel.h:12-22)."""

    def __init__(self, material):
        self.m_material = material
        self.m_material.init()          # ComputeModel.cpp:11-14

    def getMaterial(self):
        return self.m_material

    def computeCommand(self, commandBuffer, currentFrame, x, y, z):
        """bind + vkCmdDispatch(x, y, z) (ComputeModel.cpp:21-25) followed by the target -> accumulation copy the
        reference records right after it (main.cpp:253-261).  Asynchronous, like a recorded command."""
        m = self.m_material
        m.bind(commandBuffer, currentFrame)
        m._check(N.lib().vcrt_dispatch(m._ctx, int(x), int(y), int(z)))

    def frameCommand(self, commandBuffer, currentFrame, x, y, z, out=None):
        """computeCommand as a frame in flight (vcrt_frame_dispatch; between framesBegin and framesEnd): returns the slot whose
        fence (frameWait) guards `out`, a PinnedFrame or (H, W, 4) uint8 array that receives the presented rgba8 frame."""
        m = self.m_material
        m.bind(commandBuffer, currentFrame)
        arr = out.array if isinstance(out, PinnedFrame) else out
        slot = C.c_uint32()
        m._check(N.lib().vcrt_frame_dispatch(m._ctx, int(x), int(y), int(z), arr.ctypes.data if arr is not None else None, arr.nbytes if arr is not None else 0,
                                             C.byref(slot)))
        return slot.value

    def renderCommand(self, commandBuffer, currentFrame, params):
        """The same path with run-time parameters (vcrt_render): sample loop, depth, RNG/accumulation mode, sharding."""
        m = self.m_material
        m.bind(commandBuffer, currentFrame)
        if params.shader == 0 and m.m_computeShaderPath.split("/")[-1].split(".")[0] == "ray-trace-compute-simple":
            params.shader = 1
        m._check(N.lib().vcrt_render(m._ctx, C.byref(params)))
