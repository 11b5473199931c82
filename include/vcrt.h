/*
 * vcrt.h -- C ABI of the B200-native replacement for the reference's one hot path:
 * the per-pixel path-tracing compute shader resources/shaders/source/ray-trace-compute.comp
 * (and its -simple variant) of grigoryoskin/vulkan-compute-ray-tracing.
 *
 * Everything a host needs in order to do what src/main.cpp does with
 * mcvkp::ComputeMaterial / mcvkp::ComputeModel is here: plain pointers and sizes, no C++,
 * no torch types.  Each entry point cites the reference interface it replaces
 * (file:line relative to the reference tree).
 *
 * Error convention (replaces `throw std::runtime_error("failed to ...")`,
 * ComputeMaterial.cpp:35-38, Buffer.h:71-79): every call returns VCRT_OK (0) or a negative
 * code and never throws; the message is available from vcrt_last_error().
 *
 * Threading (reference: one host thread, one queue, main.cpp:323-395): one ctx = one device +
 * one CUDA stream.  Calls on one ctx must not race; different ctxs are independent.
 */
#ifndef VCRT_H
#define VCRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Data ABI: bit-identical to GpuModel::* (src/ray-tracing/GpuModels.h:18-63) and to the std430
 * mirrors in resources/shaders/source/include/definitions.glsl:1-29.
 * ---------------------------------------------------------------------------------------- */
enum { VCRT_MAT_LIGHT = 0, VCRT_MAT_LAMBERTIAN = 1, VCRT_MAT_METAL = 2, VCRT_MAT_GLASS = 3 }; /* GpuModels.h:18-24 */

typedef struct { uint32_t type; uint32_t _pad0[3]; float albedo[3]; uint32_t _pad1; } vcrt_material;        /* 32 B, GpuModels.h:26-30 */
typedef struct { float v0[3]; uint32_t _pad0; float v1[3]; uint32_t _pad1; float v2[3]; uint32_t materialIndex; } vcrt_triangle; /* 48 B, :32-38 */
typedef struct { float s[4]; uint32_t materialIndex; uint32_t _pad[3]; } vcrt_sphere;                        /* 32 B, :40-44 */
typedef struct { float min[3]; uint32_t _pad0; float max[3]; int32_t leftNodeIndex; int32_t rightNodeIndex;
                 int32_t objectIndex; uint32_t _pad1[2]; } vcrt_bvh_node;                                    /* 48 B, :47-54 */
typedef struct { uint32_t triangleIndex; float area; } vcrt_light;                                            /* 8 B,  :57-63 */
/* std140 uniform block, ray-trace-compute.comp:8-15 == UniformBufferObject, main.cpp:39-47 */
typedef struct { float camPos[3]; float time; uint32_t currentSample; uint32_t numTriangles; uint32_t numLights; uint32_t numSpheres; } vcrt_ubo; /* 32 B */

/* Descriptor bindings of the shader (ray-trace-compute.comp:8-39; order built by main.cpp:145-152) */
enum {
    VCRT_BINDING_UBO = 0, VCRT_BINDING_TARGET = 1, VCRT_BINDING_ACCUM = 2, VCRT_BINDING_TRIANGLES = 3,
    VCRT_BINDING_MATERIALS = 4, VCRT_BINDING_BVH = 5, VCRT_BINDING_LIGHTS = 6, VCRT_BINDING_SPHERES = 7
};

/* ------------------------------------------------------------------------------------------
 * Render parameters: the reference's compile-time constants turned into run-time fields.
 * Zero means "the shader's own value".
 * ---------------------------------------------------------------------------------------- */
enum { VCRT_SHADER_FULL = 0,    /* ray-trace-compute.comp        (main.cpp:144) */
       VCRT_SHADER_SIMPLE = 1   /* ray-trace-compute-simple.comp (main.cpp:143) */ };
enum { VCRT_TRAVERSAL_REFERENCE = 0, /* hit_bvh exactly as written (ray-trace-compute.comp:263-311), on the bound bvh[] */
       VCRT_TRAVERSAL_FAST = 1,      /* same closest hit (same tie rule), repacked nodes, ordered + t-culled, persistent warps */
       VCRT_TRAVERSAL_BRUTE_FORCE = 2/* hit_scene (:222-247): all triangles, then all spheres; the only mode that sees spheres */ };
enum { VCRT_RNG_PCG_REF = 0,    /* random.glsl:4-22, seed (600*x+y)*(sample+1) */
       VCRT_RNG_PHILOX = 1      /* Philox4x32-10, key (pixel, seed), counter (sample, bounce, block-in-bounce, 0) */ };
enum { VCRT_ACCUM_RGBA8_REF = 0,/* running mean through the rgba8 target/accumulation pair (ray-trace-compute.comp:375-379 + main.cpp:253-261) */
       VCRT_ACCUM_F32 = 1       /* sum of unclamped samples into an RGBA f32 buffer (A counts samples) */ };
enum { VCRT_TRIG_LIBM = 0,      /* platform sinf/cosf (glibc on the CPU, CUDA libdevice on the GPU): ulp-level differences */
       VCRT_TRIG_PORTABLE = 1   /* a fixed fp32 operation sequence shared by oracle and kernels: bit-exact across both */ };

enum { VCRT_FLAG_REF_DISPATCH_COVERAGE = 1u, /* only floor(W/32)*32 x floor(H/32)*32 pixels, as main.cpp:228 dispatches */
       VCRT_FLAG_WRITE_AOV = 2u,             /* primary-hit AOV (vcrt_aov per pixel) */
       VCRT_FLAG_COUNT_TRAVERSAL = 4u,       /* count node/triangle fetches (slower kernel variant) */
       VCRT_FLAG_STATIC_KERNEL = 8u,         /* fast traversal in the one-thread-per-pixel launch (A/B, debugging) */
       VCRT_FLAG_MEGAKERNEL = 16u,           /* fast traversal in the persistent-warps megakernel (whole paths per lane, no barrier between bounces) */
       VCRT_FLAG_WAVEFRONT = 32u             /* fast traversal in the wavefront pipeline whatever the shape of the call */
       /* none of the three: by the shape of the call -- 1 spp and depth <= 4: one launch, one thread per pixel; 1 spp and deeper: the megakernel on
          scenes below 8192 triangles when one frame is rendered at a time, else the wavefront pipeline; more samples: the wavefront pipeline */ };

typedef struct {
    uint32_t struct_size;    /* = sizeof(vcrt_render_params) */
    uint32_t shader;         /* VCRT_SHADER_* */
    uint32_t traversal;      /* VCRT_TRAVERSAL_* */
    uint32_t rng_mode;       /* VCRT_RNG_* */
    uint32_t accum_mode;     /* VCRT_ACCUM_* */
    uint32_t trig_mode;      /* VCRT_TRIG_* */
    uint32_t max_bounces;    /* NUM_BOUNCES (:313 / simple :188); 0 = 2 (full) or 4 (simple) */
    uint32_t stack_depth;    /* MAX_STACK_DEPTH (:262) of the reference traversal; 0 = 16; <= 64 */
    uint32_t lights_length;  /* what lights.length() reports; 0 = element count of binding 6 */
    uint32_t sample_begin;   /* first value of ubo.currentSample rendered by this call */
    uint32_t sample_count;   /* samples per pixel rendered by this call; 0 = 1 */
    uint32_t tile_rank;      /* image-tile sharding: 32x32 tile k (row-major) is rendered iff */
    uint32_t tile_count;     /*   k % tile_count == tile_rank;  tile_count 0 or 1 = every tile */
    uint32_t philox_seed;    /* VCRT_RNG_PHILOX key word 1 */
    uint32_t flags;          /* VCRT_FLAG_* */
    uint32_t _reserved;
} vcrt_render_params;

typedef struct { int32_t triangle; int32_t material; float t; uint32_t backFace; } vcrt_aov; /* triangle = -1 on a miss */

typedef struct {
    uint64_t rays;           /* closest-hit queries (primary + bounce) answered since the last reset: what the shader's ray_color
                                asks for, one per sample and bounce (ray-trace-compute.comp:321-323) */
    uint64_t nodes;          /* BVH node records fetched   (only with VCRT_FLAG_COUNT_TRAVERSAL) */
    uint64_t triangles;      /* triangle records fetched   (only with VCRT_FLAG_COUNT_TRAVERSAL) */
    double   kernel_ms;      /* device time of the render kernels (CUDA events) since the last reset */
    uint64_t launches;       /* kernels launched since the last reset */
    double   trace_ms;       /* device time of the dominant kernel alone (wavefront trace launches; CUDA events per launch, summed).  With two or more
                                pipelines ("wf_streams") the launches of different pipelines overlap and each one's events also span the time it shares
                                the SMs: set wf_streams=1 to time the kernel by itself */
    uint64_t trace_launches; /* number of those launches */
    uint64_t traversals;     /* BVH traversals actually run.  The wavefront pipeline traces bounce 0 once per pixel and shares the hit
                                among the pixel's samples (the shader's primary ray does not depend on the sample, :352-373), so
                                traversals = rays - primary_rays * (1 - 1/sample_count) there; every other kernel: = rays */
    uint64_t primary_rays;   /* the bounce-0 part of `rays` */
    double   primary_trace_ms; /* the part of trace_ms spent in bounce-0 launches */
} vcrt_counters;

enum { VCRT_OK = 0, VCRT_ERR_INVALID = -1, VCRT_ERR_CUDA = -2, VCRT_ERR_STATE = -3, VCRT_ERR_NOMEM = -4 };

typedef struct vcrt_ctx vcrt_ctx;

/* Context = device + stream + every device allocation.  Replaces VulkanApplicationContext bring-up
 * (VulkanApplicationContext.cpp:47-119) + ComputeMaterial's pipeline/descriptor objects (ComputeMaterial.cpp:15-61). */
int vcrt_create(int device, vcrt_ctx** out);
int vcrt_destroy(vcrt_ctx* ctx);                     /* ~Material, Material.cpp:21-29; ~Buffer, Buffer.h:21-30 */
const char* vcrt_last_error(const vcrt_ctx* ctx);    /* ctx may be NULL: error of the last failed vcrt_create on this thread */
const char* vcrt_version(void);

/* Selects the kernel by the reference's shader name ("ray-trace-compute" / "ray-trace-compute-simple", with or
 * without directory and .spv/.comp suffix): the ComputeMaterial(computeShaderPath) argument, ComputeMaterial.cpp:9-13. */
int vcrt_set_shader(vcrt_ctx* ctx, const char* shader_path);

/* Storage buffers, bindings 3..7 (BufferUtils::createBundle<T> + addStorageBufferBundle, main.cpp:88-106, :148-152).
 * `host` is copied; bytes must be a multiple of the record size.  Element counts come from byte sizes. */
int vcrt_set_buffer(vcrt_ctx* ctx, int binding, const void* host, size_t bytes);
/* Same, from a device pointer on ctx's device (copied device-to-device on ctx's stream). */
int vcrt_set_buffer_device(vcrt_ctx* ctx, int binding, const void* dev, size_t bytes);

/* Storage images, bindings 1 and 2: both rgba8 W x H (ImageUtils::createImage, main.cpp:108-140), zero-filled.
 * Also (re)allocates the f32 accumulation buffer and the AOV buffer. */
int vcrt_set_image_size(vcrt_ctx* ctx, uint32_t width, uint32_t height);

/* The per-frame 32-byte uniform write (updateScene, main.cpp:166-183). */
int vcrt_set_ubo(vcrt_ctx* ctx, const vcrt_ubo* ubo);

/* ComputeModel::computeCommand(cmd, frame, x, y, z) == bind + vkCmdDispatch(x, y, z) (ComputeModel.cpp:21-25), followed
 * by the target -> accumulation image copy the reference records right after it (main.cpp:253-261).  One sample
 * (ubo.currentSample), reference traversal, reference RNG, rgba8 running mean, x*32 by y*32 pixels. Asynchronous. */
int vcrt_dispatch(vcrt_ctx* ctx, uint32_t groups_x, uint32_t groups_y, uint32_t groups_z);

/* The same path with the shader's compile-time constants as parameters and the sample loop inside the kernel.
 * camPos is taken from the current UBO; ubo.currentSample is ignored (params.sample_begin is used). Asynchronous. */
int vcrt_render(vcrt_ctx* ctx, const vcrt_render_params* params);

int vcrt_clear_accum(vcrt_ctx* ctx);                                /* zero target, accumulation, f32 accumulation, AOV */
int vcrt_resolve(vcrt_ctx* ctx, uint32_t total_samples, float gamma); /* f32 accumulation / total -> clamp -> pow(1/gamma) -> target rgba8 (gamma<=0: none; 2.2 = post-process-shader.frag:67-68) */

/* The reference's post-process pass (resources/shaders/source/post-process-shader.frag:26-70) on the rgba8 target, into a
 * ctx-owned rgba8 "present" image (what the reference draws to the swapchain): out.rgb = pow(mix * smartDeNoise(target, sigma,
 * kSigma, threshold) + (1 - mix) * target, 1/gamma), out.a = 1.  mix = 0 and gamma = 2.2 is the shipped shader (the denoiser
 * call is commented out, :64-65, where it would run with mix 0.5, sigma 2, kSigma 2, threshold 0.05); gamma <= 0: none. */
int vcrt_post_process(vcrt_ctx* ctx, float mix, float sigma, float k_sigma, float threshold, float gamma);
int vcrt_read_present_rgba8(vcrt_ctx* ctx, void* dst, size_t bytes);

/* Read-backs (synchronise ctx's stream).  bytes must equal the buffer size. */
int vcrt_read_target_rgba8(vcrt_ctx* ctx, void* dst, size_t bytes); /* getStorageImages()[0], main.cpp:194 */
int vcrt_read_accum_rgba8(vcrt_ctx* ctx, void* dst, size_t bytes);  /* getStorageImages()[1], main.cpp:195 */
int vcrt_read_accum_f32(vcrt_ctx* ctx, void* dst, size_t bytes);    /* W*H*4 floats */
int vcrt_read_aov(vcrt_ctx* ctx, void* dst, size_t bytes);          /* W*H vcrt_aov */
int vcrt_write_accum_f32(vcrt_ctx* ctx, const void* src, size_t bytes); /* resume: reload a dumped accumulation */

/* Device pointers of ctx-owned images, for collectives (NCCL) and zero-copy consumers. */
int vcrt_device_ptr(vcrt_ctx* ctx, int what /* 0 target rgba8, 1 accum rgba8, 2 accum f32, 3 aov */, void** out, size_t* bytes);

/* Multi-GPU tile sharding (no reference counterpart: one VkDevice, VulkanApplicationContext.cpp:95-119).  Copies the
 * pixels of the 32x32 tiles k with k % tile_count == tile_rank (row-major tile order, the partition of
 * vcrt_render_params.tile_rank/tile_count) between a ctx image (what: 0 target rgba8, 2 f32 accumulation) and a packed
 * DEVICE buffer: owned tiles back to back, 1024 pixels each, row-major inside the tile, zero outside the image.
 * bytes >= owned_tiles * 1024 * (4 | 16).  Asynchronous on ctx's stream; pair with an all-gather of the packed buffers. */
int vcrt_pack_tiles(vcrt_ctx* ctx, int what, uint32_t tile_rank, uint32_t tile_count, void* packed_dev, size_t bytes);
int vcrt_unpack_tiles(vcrt_ctx* ctx, int what, uint32_t tile_rank, uint32_t tile_count, const void* packed_dev, size_t bytes);

/* Tunables that do not change results.  "fast_bvh": "sah" (default; the fast traversal walks a surface-area-heuristic
 * tree built over the leaves of the bound bvh[], optimised by reinsertion for scenes up to 2 Mi triangles), "sah_plain" (the same tree
 * as built, a tenth of the host time) or "topology" (it keeps the bound tree's own topology); "fast_nodes": "auto"
 * (default: 4-wide quantised 64-byte nodes when the scene extent allows, else binary 64-byte float nodes), "q15x4" (4-wide quantised
 * whatever the extent), "q15" (binary quantised 32-byte nodes), "f32"; "fast_build": where the records are built -- "auto" (default: on the
 * device, by CUDA kernels reading the bound buffers where they lie, whenever the default tree is wanted (fast_bvh=sah with fast_nodes=auto|q15x4)
 * and the scene allows; else by the host builder), "host", "device" (fail instead of falling back); "dispatch_traversal": what vcrt_dispatch walks -- "auto" (default: the fast tree whenever
 * the bound tree is at most 13 levels deep, i.e. whenever the shader's 16-entry stack cannot overflow; identical frames), "reference" (always the
 * literal hit_bvh), "fast"; "wf_batch_paths":
 *  paths per wavefront batch (queue memory: 120 B per path); "wf_streams": "auto" (default) or 1..4 -- the render is cut
 * into that many batches, run as parallel pipelines on separate streams with their own queue sets, which share the queue memory of one
 * batch; "auto" = 2 for calls of 32 Mi paths or more (one pipeline's shade launches overlap the other's trace launches: 3-5 % faster
 * on a 64-spp 1080p frame), else 1 (slower than one pipeline on small frames); frames in flight run one pipeline each;
 * vcrt_get_info "wf_pipelines" = what the last render used; "trace_timing": "auto" (default: multi-sample renders only) | "on" | "off" -- CUDA events around every trace launch (vcrt_counters.trace_ms);
 * "leaf_threshold" / "shade_threshold" / "continue_threshold": lanes (1..32); "host_threads": OpenMP threads of the host-side record build. */
int vcrt_set_option(vcrt_ctx* ctx, const char* key, const char* value);

/* Read-only facts about ctx as text: "fast_nodes" -> "q15x4" | "q15" | "f32" | "none" (what the fast traversal walks after the last
 * upload), "fast_node_count", "fast_depth", "wf_batch_paths", "dispatch_kernel" ("fast" | "reference": what the last vcrt_dispatch walked), "fast_build" ("device" | "host": where the current records were
 * built), "fast_build_ms" (wall time of that build), "device".  Builds the fast records if they are stale. */
int vcrt_get_info(vcrt_ctx* ctx, const char* key, char* value, size_t capacity);

/* Run ctx's work on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL restores ctx's own stream), so
 * that renders order with the caller's collectives / events without extra synchronisation.  Synchronises first. */
int vcrt_set_stream(vcrt_ctx* ctx, void* cuda_stream);

int vcrt_synchronize(vcrt_ctx* ctx);
int vcrt_get_counters(vcrt_ctx* ctx, vcrt_counters* out);           /* synchronises */
int vcrt_reset_counters(vcrt_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Frames in flight.  The reference's frame loop keeps MAX_FRAMES_IN_FLIGHT = 2 frames going (main.cpp:68; fences and semaphores
 * :298-316; drawFrame waits for the fence of the slot it is about to reuse, :325, submits, :367-380, and moves on, :394), so the
 * "ms/frame" it prints (:397-413) is the rate of a pipelined loop.  The same here: between vcrt_frames_begin and vcrt_frames_end,
 * vcrt_frame_submit renders ONE sample per pixel on the slot's own stream -- the render kernels of consecutive frames overlap, the
 * long rays that end one frame run beside the bulk of the next --, folds it into the accumulation IN FRAME ORDER (an event chains the
 * folds), resolves, and copies the rgba8 frame into `host_dst` (pinned memory: vcrt_alloc_host) while later frames render.  Frames
 * are bit-identical to the synchronous sequence vcrt_set_ubo / vcrt_render / vcrt_resolve / vcrt_read_target_rgba8.
 *   params       as for vcrt_render, sample_count 0 or 1; VCRT_ACCUM_RGBA8_REF: the running mean through the rgba8 pair;
 *                VCRT_ACCUM_F32: the f32 sum, which restarts when params->sample_begin == 0 (as the running mean does: its weight
 *                of the history is 0 at sample 0, ray-trace-compute.comp:375-379) and is resolved with 1/total_samples
 *                (0 = sample_begin + 1) and `gamma` as by vcrt_resolve
 *   host_dst     W*H*4 bytes or NULL (no read-back); valid after vcrt_frame_wait(slot) -- the vkWaitForFences of that slot -- and
 *                not to be freed or reused before that (the copy into it is asynchronous)
 * While frames are in flight the other calls that touch the context's images or buffers fail with VCRT_ERR_STATE; vcrt_set_ubo,
 * vcrt_set_option (thresholds), vcrt_get_info and vcrt_last_error are fine.  After vcrt_frames_end the target, the accumulation and
 * the counters are those of the last frame, as if the frames had been rendered synchronously.
 * ---------------------------------------------------------------------------------------- */
#define VCRT_MAX_FRAMES_IN_FLIGHT 4
int vcrt_frames_begin(vcrt_ctx* ctx, uint32_t frames_in_flight /* 1..4; the reference: 2 */);
int vcrt_frame_submit(vcrt_ctx* ctx, const vcrt_render_params* params, uint32_t total_samples, float gamma, void* host_dst, size_t bytes, uint32_t* slot);
/* ComputeModel::computeCommand (vcrt_dispatch: ubo.currentSample, reference RNG, rgba8 running mean, x*32 by y*32 pixels) as a frame in flight */
int vcrt_frame_dispatch(vcrt_ctx* ctx, uint32_t groups_x, uint32_t groups_y, uint32_t groups_z, void* host_dst, size_t bytes, uint32_t* slot);
int vcrt_frame_wait(vcrt_ctx* ctx, uint32_t slot);
int vcrt_frames_end(vcrt_ctx* ctx);
/* Page-locked host memory for read-backs that overlap rendering (the reference's staging buffers are host-visible VMA
 * allocations, Buffer.h:42-60).  Any host memory works with every call; only pinned memory keeps vcrt_frame_submit asynchronous. */
int vcrt_alloc_host(size_t bytes, void** out);
int vcrt_free_host(void* ptr);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU groups (no reference counterpart: one VkDevice and one queue, VulkanApplicationContext.cpp:95-119; the shader's
 * invocations never communicate, so pixels and samples partition freely -- SURVEY.md 8e).  The scene is replicated; one call
 * renders one frame on all GPUs of the group and combines the shares with ONE NCCL collective on the render streams:
 *   VCRT_SHARD_TILES    32x32 tile k (row-major) is rendered by rank k % world; every rank resolves its tiles to rgba8, the
 *                       packed tiles are all-gathered and scattered: EVERY rank ends up with the full rgba8 target, bit-identical
 *                       to a 1-GPU render (both RNG modes are keyed by the pixel)
 *   VCRT_SHARD_SAMPLES  rank r renders a contiguous slice of the samples of every pixel; the f32 accumulation buffers are
 *                       sum-reduced onto rank 0, which resolves: RANK 0 ends up with the full target (same sample set as a 1-GPU
 *                       render; fp32 summation order differs)
 * Two ways to form a group: one process driving n devices (create_local: the group owns n contexts, the set_* calls below
 * broadcast to them), or one process per GPU (create_rank: wraps the caller's context; rank 0 obtains an id with
 * vcrt_group_unique_id and the caller distributes its VCRT_GROUP_ID_BYTES bytes, e.g. through torch.distributed / MPI).
 * NCCL (libnccl.so.2) is loaded at run time; groups of size 1 need none.
 * ---------------------------------------------------------------------------------------- */
typedef struct vcrt_group vcrt_group;
enum { VCRT_SHARD_TILES = 0, VCRT_SHARD_SAMPLES = 1 };
#define VCRT_GROUP_ID_BYTES 128

int vcrt_group_unique_id(void* id /* VCRT_GROUP_ID_BYTES */);
int vcrt_group_create_local(int n_devices, const int* devices /* NULL = 0..n-1 */, vcrt_group** out);
int vcrt_group_create_rank(vcrt_ctx* ctx /* borrowed */, const void* id, int rank, int world, vcrt_group** out);
int vcrt_group_destroy(vcrt_group* group);                        /* a local group destroys its contexts too */
const char* vcrt_group_last_error(const vcrt_group* group);       /* group may be NULL: error of the last failed create on this thread */
int vcrt_group_size(const vcrt_group* group);                     /* ranks in the group */
int vcrt_group_local_count(const vcrt_group* group);              /* contexts this process drives */
int vcrt_group_rank(const vcrt_group* group, int local_index);    /* rank of local context i */
vcrt_ctx* vcrt_group_ctx(vcrt_group* group, int local_index);     /* for per-context calls (options, counters, read-backs) */
/* the single-context setters, applied to every local context (main.cpp:84-153 once per GPU) */
int vcrt_group_set_shader(vcrt_group* group, const char* shader_path);
int vcrt_group_set_buffer(vcrt_group* group, int binding, const void* host, size_t bytes);
int vcrt_group_set_image_size(vcrt_group* group, uint32_t width, uint32_t height);
int vcrt_group_set_ubo(vcrt_group* group, const vcrt_ubo* ubo);
int vcrt_group_set_option(vcrt_group* group, const char* key, const char* value);
/* One frame: clears the accumulation, renders params (accum_mode must be VCRT_ACCUM_F32; sample_count = the frame's total;
 * tile_rank/tile_count must be 0) sharded by `mode`, combines, resolves with `gamma` (vcrt_resolve).  Asynchronous. */
int vcrt_group_render(vcrt_group* group, const vcrt_render_params* params, int mode, float gamma);
int vcrt_group_read_target_rgba8(vcrt_group* group, int local_index, void* dst, size_t bytes);   /* synchronises that context */
int vcrt_group_synchronize(vcrt_group* group);

#ifdef __cplusplus
}
#endif
#endif /* VCRT_H */
