"""B200-native replacement for the hot path of grigoryoskin/vulkan-compute-ray-tracing: the per-pixel path-tracing
compute shader, behind the reference's ComputeMaterial / ComputeModel interface.  CUDA only -- no CPU path."""
from .api import (ACCUM, AOV_DTYPE, FLAG_COUNT_TRAVERSAL, FLAG_MEGAKERNEL, FLAG_STATIC_KERNEL, FLAG_WAVEFRONT, FLAG_REF_DISPATCH_COVERAGE, FLAG_WRITE_AOV, RNG, SHADER, TRAVERSAL, TRIG,
                  VK_SHADER_STAGE_COMPUTE_BIT, Buffer, BufferBundle, BufferUtils, ComputeMaterial, ComputeModel, Image, PinnedFrame,
                  VcrtError, render_params)
from . import imageio
from .frameloop import Camera, FrameLoop
from .scene import CAMERA_START, load_scene, pack_ubo, save_scene

__all__ = [n for n in dir() if not n.startswith("_")]
