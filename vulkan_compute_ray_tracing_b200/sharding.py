"""Multi-GPU sharding of the hot path: one process per GPU, the scene replicated, one exchange step at the end.

The reference has no multi-device path (one VkDevice, one queue: VulkanApplicationContext.cpp:95-119); the shader's
invocations never communicate (ray-trace-compute.comp has no shared memory, atomics or barriers), so pixels and samples
partition freely.  Two partitions, both expressed through fields of vcrt_render_params:

  "tiles"    32x32 tiles (the reference's workgroup footprint, ray-trace-compute.comp:3) in row-major order, tile k on
             rank k % world.  Every pixel is rendered by exactly one rank with the RNG stream it has in a 1-GPU run
             (both RNG modes are keyed by pixel position), so the combined frame is bit-identical to the 1-GPU frame.
             Exchange: packed owned tiles -> all_gather -> unpack (compact, equal counts per rank), or a SUM reduce of
             the full-size buffers (a pixel's other contributions are +0.0, which leaves its bits unchanged).
  "samples"  rank r renders samples [begin_r, begin_r + count_r) of EVERY pixel into its f32 sum buffer; exchange: SUM
             reduce.  The sample set equals the 1-GPU run's (seeds depend on the global sample index, random.glsl:19 /
             the Philox counter); only the fp32 summation order differs.

On the GPUs the partition AND the exchange live in the C ABI (include/vcrt.h: vcrt_group_*, NCCL inside libvcrt.so on the
render streams); `Group` / `LocalGroup` below are thin callers.  torch.distributed only carries the 128-byte NCCL id between
the processes of a torchrun launch.  The pure-host functions (tile/sample partition, packed-tile layout, `render_sharded`
over any torch.distributed backend) define the same partition for the gloo CPU tests and for checking the kernels.
"""
import copy

import numpy as np

TILE = 32


def tile_grid(width, height):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def owned_tiles(width, height, rank, world):
    """Row-major tile indices rendered by `rank` (k % world == rank) -- the enumeration of item_to_pixel (vcrt_path.cuh)."""
    tx, ty = tile_grid(width, height)
    return np.arange(rank, tx * ty, max(world, 1), dtype=np.int64)


def max_owned_tiles(width, height, world):
    tx, ty = tile_grid(width, height)
    return (tx * ty + world - 1) // world


def sample_slices(total_samples, world, first_sample=0):
    """Contiguous sample ranges, one per rank; the first `total % world` ranks take one extra sample."""
    base, extra = divmod(int(total_samples), world)
    out, begin = [], int(first_sample)
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((begin, n))
        begin += n
    return out


def shard_params(params, mode, rank, world, total_samples=None):
    """A copy of `params` restricted to this rank's share.  Returns (params, renders_anything)."""
    p = copy.copy(params)
    if world <= 1:
        return p, True
    if mode == "tiles":
        p.tile_rank, p.tile_count = rank, world
        return p, True
    if mode == "samples":
        total = int(total_samples if total_samples is not None else (params.sample_count or 1))
        begin, n = sample_slices(total, world, params.sample_begin)[rank]
        p.sample_begin, p.sample_count = begin, max(n, 1)
        return p, n > 0       # sample_count 0 means "1" to the library, so an empty slice must not be launched
    raise ValueError("sharding mode must be 'tiles' or 'samples'")


def tile_pixel_index(width, height, rank, world, pad_tiles=None):
    """Flat pixel indices (y * W + x) of the packed layout: owned tiles in order, 1024 pixels each in row-major order
    inside the tile; pixels outside the image (ragged edge tiles) and padding tiles are -1.  This is the layout of
    vcrt_pack_tiles / vcrt_unpack_tiles."""
    tx, _ = tile_grid(width, height)
    tiles = owned_tiles(width, height, rank, world)
    n = len(tiles) if pad_tiles is None else int(pad_tiles)
    idx = np.full((n, TILE, TILE), -1, np.int64)
    if len(tiles):
        t_y, t_x = np.divmod(tiles, tx)
        ys = t_y[:, None, None] * TILE + np.arange(TILE)[None, :, None]
        xs = t_x[:, None, None] * TILE + np.arange(TILE)[None, None, :]
        ok = (ys < height) & (xs < width)
        idx[: len(tiles)] = np.where(ok, ys * width + xs, -1)
    return idx.reshape(-1)


def reduce_accumulation(accum, group=None, dst=0):
    """SUM-reduce of the f32 accumulation buffers onto `dst` (in place on dst).  Valid for both partitions."""
    import torch.distributed as dist
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def gather_tiles(packed, width, height, world, out_flat, group=None):
    """all_gather of equally sized packed tile buffers ((max_owned_tiles*1024, C) per rank) and scatter of every rank's
    pixels into out_flat ((W*H, C)).  Runs on whatever device the tensors live on."""
    import torch
    import torch.distributed as dist
    n = max_owned_tiles(width, height, world)
    assert packed.shape[0] == n * TILE * TILE, "packed buffer must be padded to max_owned_tiles"
    gathered = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed.contiguous(), group=group)
    gathered = gathered.view((world,) + tuple(packed.shape))
    for r in range(world):
        idx = torch.from_numpy(tile_pixel_index(width, height, r, world, pad_tiles=n)).to(packed.device)
        keep = idx >= 0
        out_flat[idx[keep]] = gathered[r][keep]
    return out_flat


def gather_tiles_device(material, what, rank, world, group=None):
    """Tile-sharded frame -> complete frame on every rank, on the GPU: vcrt_pack_tiles (CUDA) -> NCCL all_gather of the
    equally sized packed buffers -> vcrt_unpack_tiles (CUDA) of every peer's tiles into this rank's image.
    what: 0 rgba8 target, 2 f32 accumulation.  Everything is enqueued on torch's current stream, which must be the
    stream the material renders on (ComputeMaterial.setStream)."""
    import torch
    import torch.distributed as dist
    img = material.getStorageImages()[0].data
    elem = 4 if what == 0 else 16
    n = max_owned_tiles(img.width, img.height, world) * TILE * TILE * elem
    mine = torch.empty(n, dtype=torch.uint8, device="cuda")
    material.packTiles(what, rank, world, mine.data_ptr(), n)
    if world == 1:
        return
    everyone = torch.empty(world * n, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(everyone, mine, group=group)
    for r in range(world):
        if r != rank:
            material.unpackTiles(what, r, world, everyone.data_ptr() + r * n, n)


def render_sharded(render, params, mode, rank, world, total_samples=None, group=None, dst=0):
    """render(params) must add this rank's samples into its accumulation tensor ((H, W, 4) f32, zeroed by the caller
    beforehand) and return it.  After the call rank `dst` holds the combined sums."""
    p, active = shard_params(params, mode, rank, world, total_samples)
    accum = render(p if active else None)
    if world > 1:
        reduce_accumulation(accum, group=group, dst=dst)
    return accum


MODES = {"tiles": 0, "samples": 1}     # VCRT_SHARD_TILES / VCRT_SHARD_SAMPLES


class Group:
    """One process per GPU (torchrun): vcrt_group_create_rank around this process's ComputeMaterial.  `render` = one frame on
    all GPUs: this rank's share + ONE NCCL collective inside libvcrt.so on the render stream + resolve."""

    def __init__(self, material, rank, world, id_bytes=None):
        import ctypes as C
        from . import _native as N
        material.init()
        self.material, self.rank, self.world = material, rank, world
        self._g = C.c_void_p()
        L = N.lib()
        buf = C.create_string_buffer(bytes(id_bytes), 128) if id_bytes is not None else None
        if L.vcrt_group_create_rank(material._ctx, buf, rank, world, C.byref(self._g)) != 0:
            raise N.VcrtError(L.vcrt_group_last_error(None).decode())

    @staticmethod
    def unique_id():
        import ctypes as C
        from . import _native as N
        buf = C.create_string_buffer(128)
        if N.lib().vcrt_group_unique_id(buf) != 0:
            raise N.VcrtError(N.lib().vcrt_group_last_error(None).decode())
        return buf.raw

    @classmethod
    def from_torch(cls, material, group=None):
        """Rank 0 creates the NCCL id; torch.distributed (any backend) hands it to the other ranks."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 and world > 1 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        return cls(material, rank, world, box[0])

    def render(self, model, params, mode, gamma=0.0, currentFrame=0):
        import ctypes as C
        from . import _native as N
        model.getMaterial().bind(None, currentFrame)        # the frame's UBO, as computeCommand does
        if N.lib().vcrt_group_render(self._g, C.byref(params), MODES[mode], float(gamma)) != 0:
            raise N.VcrtError(N.lib().vcrt_group_last_error(self._g).decode())

    def close(self):
        from . import _native as N
        if self._g is not None:
            N.lib().vcrt_group_destroy(self._g)
            self._g = None


class LocalGroup:
    """One process driving n GPUs: vcrt_group_create_local; the scene goes to every GPU through the group's setters."""

    def __init__(self, n_devices, scene, width, height, shader="ray-trace-compute", devices=None):
        import ctypes as C
        from . import _native as N
        L = N.lib()
        self._g = C.c_void_p()
        devs = (C.c_int * n_devices)(*devices) if devices is not None else None
        if L.vcrt_group_create_local(n_devices, devs, C.byref(self._g)) != 0:
            raise N.VcrtError(L.vcrt_group_last_error(None).decode())
        self.n, self.width, self.height = n_devices, width, height
        self._check(L.vcrt_group_set_shader(self._g, shader.encode()))
        self._check(L.vcrt_group_set_image_size(self._g, width, height))
        for i, name in enumerate(("triangles", "materials", "bvh", "lights", "spheres")):
            a = np.ascontiguousarray(scene[name]).view(np.uint8).reshape(-1)
            self._check(L.vcrt_group_set_buffer(self._g, 3 + i, a.ctypes.data if a.nbytes else None, a.nbytes))

    def _check(self, rc):
        from . import _native as N
        if rc != 0:
            raise N.VcrtError(N.lib().vcrt_group_last_error(self._g).decode())

    def render(self, ubo_bytes, params, mode, gamma=0.0):
        import ctypes as C
        from . import _native as N
        ubo = N.Ubo.from_buffer_copy(bytes(ubo_bytes))
        self._check(N.lib().vcrt_group_set_ubo(self._g, C.byref(ubo)))
        self._check(N.lib().vcrt_group_render(self._g, C.byref(params), MODES[mode], float(gamma)))

    def read_target(self, local_index=0):
        from . import _native as N
        out = np.empty((self.height, self.width, 4), np.uint8)
        self._check(N.lib().vcrt_group_read_target_rgba8(self._g, local_index, out.ctypes.data, out.nbytes))
        return out

    def close(self):
        from . import _native as N
        if self._g is not None:
            N.lib().vcrt_group_destroy(self._g)
            self._g = None
