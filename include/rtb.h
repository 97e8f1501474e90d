/*
 * rtb.h — C ABI of librtb.so, the B200 (sm_100a) replacement for the reference's
 * per-pixel sample job.
 *
 * What it replaces (paths relative to /root/reference/RaytracingInOneWeekend/Assets/Scripts):
 *   Runtime/Jobs/SampleBatchJob.cs:17-164   struct SampleBatchJob : IJobParallelFor   (the job)
 *   Unity/Raytracer.cs:671-736              the host fills the job's fields and schedules it
 * The reference has no FFI for this path (it is a Burst job).  Its only native-call
 * precedent — and the style this header follows — is the denoiser binding:
 *   Assets/ThirdParty/nVidia OptiX Denoiser/OptixApi.cs:154-251   [DllImport] externs on IntPtr handles
 *   OptixDenoiser/OptixDenoiser/OptixDenoiser.h:1-10              extern "C" exports returning an int code
 *   Runtime/Jobs/DenoiseJobs.cs:10-39                             a non-Burst IJob making the blocking call
 * The P/Invoke stub a maintainer would add is in INTEGRATION.md and
 * raytracing-in-one-weekend_b200/bindings/B200PathTracerApi.cs.
 *
 * Conventions: every call returns 0 on success or an rtb_status code; never throws,
 * never aborts.  Only PODs and plain pointers cross the boundary.  One call in flight
 * per context (the host serialises batches through job dependencies,
 * Raytracer.cs:809-811); callable from any thread.  Host pointers are borrowed for
 * the duration of the call only (DenoiseJobs.cs:26-37 precedent).
 *
 * Pixel layout is the reference's: index = row * W + col, row 0 = bottom
 * (SampleBatchJob.cs:64-67).  float3 arrays are packed 12-byte elements.
 */
#ifndef RTB_H
#define RTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RTB_API __declspec(dllexport)
#else
#define RTB_API __attribute__((visibility("default")))
#endif

#define RTB_ABI_VERSION 2

typedef enum rtb_status {
  RTB_OK = 0,
  RTB_ERR_INVALID_ARGUMENT = 1,
  RTB_ERR_NO_SCENE = 2,
  RTB_ERR_CANCELLED = 3,        /* the cancellation token was set; outputs unspecified (SampleBatchJob.cs:61) */
  RTB_ERR_UNSUPPORTED = 4,      /* entity/material/sky type outside the supported hot path */
  RTB_ERR_OUT_OF_MEMORY = 5,
  RTB_ERR_CUDA = 100            /* 100 + cudaError_t */
} rtb_status;

/* ---- scene (uploaded once per world change; Raytracer.cs:1167-1183) ------------------- */

/* Entity of EntityType.Sphere (Entity.cs:13-20, EntityTypes/Sphere.cs:6-24) with a static
 * rigid transform.  `radius` may be NEGATIVE: the intersection uses radius^2 and the normal
 * is point/radius, so a negative radius flips the normal (hollow glass; HitTests.cs:43).
 * The reference keeps a quaternion per entity; a sphere is rotation-invariant, so the
 * host's flattener passes only the translation (Entity.OriginTransform.pos). */
typedef struct rtb_sphere {
  float center[3];
  float radius;
  uint32_t material;            /* index into the material array (Entity.Material - materialBuffer.ptr) */
  uint32_t reserved[3];
} rtb_sphere;                   /* 32 bytes */

/* Entity of EntityType.Triangle (Entity.cs:13-20, EntityTypes/Triangle.cs:8-29): what the host's mesh
 * ingestion produces (AddMeshRuntimeEntitiesJob.cs; Raytracer.cs:1185-1304).  Triangles are always in
 * world space (Entity.cs:92-93), so there is no transform.  Fields are the reference's: Data columns
 * (v2 - v0, v1 - v0, v0) and the three (already normalised) vertex normals. */
typedef struct rtb_triangle {
  float edge2[3];               /* Triangle.Data[0] = v2 - v0 */
  float edge1[3];               /* Triangle.Data[1] = v1 - v0 */
  float v0[3];                  /* Triangle.Data[2] */
  float normals[3][3];          /* Triangle.Normals columns: n0, n1, n2 */
  uint32_t material;
  uint32_t reserved;
} rtb_triangle;                 /* 80 bytes */

typedef enum rtb_entity_type {  /* Entity.cs:13-20 */
  RTB_ENTITY_SPHERE = 1,
  RTB_ENTITY_RECT = 2,          /* always a placed entity (rtb_upload_placed_world) */
  RTB_ENTITY_BOX = 3,           /* always a placed entity */
  RTB_ENTITY_TRIANGLE = 4,
  /* Flag ORed into rtb_entity.type: `index` points into the placed-entity array (the reference's full Entity
   * record: rotation, motion, content) instead of the sphere array.  Implied for RECT and BOX. */
  RTB_ENTITY_PLACED = 0x100
} rtb_entity_type;

/* One element of the BVH-ordered entity list the leaves point into (bvhEntities, BvhNodeData.cs:157-160):
 * the entity's type and its index in the sphere / triangle / placed-entity array (Entity.Content). */
typedef struct rtb_entity {
  uint32_t type;                /* rtb_entity_type (| RTB_ENTITY_PLACED) */
  uint32_t index;
} rtb_entity;

/* The reference's Entity (Entity.cs:24-55) with its Content inlined, for everything a plain rtb_sphere cannot
 * say: a rotated entity, a moving one (Entity.TransformAtTime, Entity.cs:124-127, evaluated at the ray's time),
 * EntityType.Rect (EntityTypes/Rect.cs: the XY rectangle of `size` centred on the entity's origin, hit from
 * +Z only, HitTests.cs:62-78) and EntityType.Box (EntityTypes/Box.cs, HitTests.cs:80-111).  The ray is taken
 * into entity space with the inverse of the transform (Entity.cs:74-103). */
typedef struct rtb_placed_entity {
  uint32_t type;                /* RTB_ENTITY_SPHERE, RTB_ENTITY_RECT or RTB_ENTITY_BOX */
  uint32_t material;
  uint32_t moving;              /* Entity.Moving */
  uint32_t reserved;
  float rotation[4];            /* OriginTransform.rot as (x, y, z, w) */
  float position[3];            /* OriginTransform.pos */
  float destination_offset[3];  /* DestinationOffset */
  float time_range[2];          /* TimeRange (moving: x != y, Entity.cs:53-54) */
  float size[3];                /* Sphere: (radius, -, -); Rect: the ctor's size.xy; Box: the ctor's size.xyz */
  float reserved2;
} rtb_placed_entity;            /* 80 bytes */

typedef enum rtb_material_type {     /* Material.cs:9-14 */
  RTB_MATERIAL_STANDARD = 0,
  RTB_MATERIAL_DIELECTRIC = 1,
  RTB_MATERIAL_PROBABILISTIC_VOLUME = 2   /* participating medium (Material.cs:48-65,163-168): albedo = Albedo.MainColor, density in
                                           * index_of_refraction (Material.parameter); worlds in which an entity wears one run the
                                           * megakernel's media flavour (RTB_OPT_KERNEL = 1, RTB_OPT_NOISE = 1 or image textures: the
                                           * collect-all kernel, bit-identical to the CPU restatement) */
} rtb_material_type;

/* Material with constant textures only (Material.cs:16-26, Texture.cs:50-59,101-108:
 * TextureType.Constant / ConstantScalar).  "Lambertian" = STANDARD{metallic 0, glossiness 0};
 * "Metal(fuzz)" = STANDARD{metallic 1, glossiness 1-fuzz}; glass = DIELECTRIC{glossiness 1}. */
typedef struct rtb_material {
  uint32_t type;                /* rtb_material_type */
  float albedo[3];              /* Albedo.MainColor */
  float emission[3];            /* Emission.MainColor */
  float glossiness;             /* Glossiness scalar */
  float metallic;               /* Metallic scalar */
  float index_of_refraction;    /* Material.parameter: IndexOfRefraction (Dielectric) or Density (ProbabilisticVolume) */
  uint32_t reserved[2];
} rtb_material;                 /* 48 bytes */

/* Flattened BvhNode (BvhNode.cs:5-22): pointers become indices into the node array
 * (root = node 0, the order BuildRuntimeBvhJob.cs:18-39 produces) and into the sphere array
 * (the reference's leaves point into a contiguous BVH-ordered entity copy,
 * BvhNodeData.cs:157-160).  Leaf iff first_entity >= 0 (BvhNode.IsLeaf: EntitiesStart != null). */
typedef struct rtb_bvh_node {
  float bounds_min[3];
  float bounds_max[3];
  int32_t left;                 /* node index or -1 */
  int32_t right;                /* node index or -1 */
  int32_t first_entity;         /* sphere index or -1 for inner nodes */
  int32_t entity_count;
} rtb_bvh_node;                 /* 40 bytes */

/* ---- per-batch uniforms (the public fields of SampleBatchJob, SampleBatchJob.cs:23-39) - */

typedef struct rtb_view {       /* View.cs:8-14, 88 bytes */
  float origin[3];
  float lower_left_corner[3];
  float horizontal[3];
  float vertical[3];
  float forward[3];
  float up[3];
  float right[3];
  float lens_radius;
} rtb_view;

typedef enum rtb_sky_type {     /* Environment.cs:5-10 */
  RTB_SKY_NONE = 0,
  RTB_SKY_GRADIENT = 1,
  RTB_SKY_CUBEMAP = 2           /* needs rtb_upload_sky_cubemap */
} rtb_sky_type;

typedef struct rtb_environment {/* Environment.cs:12-18 */
  uint32_t sky_type;
  float sky_bottom_color[3];
  float sky_top_color[3];
} rtb_environment;

typedef struct rtb_batch_params {
  float size[2];                        /* Size (float2: W, H) */
  int32_t slice_offset;                 /* SliceOffset */
  int32_t slice_divider;                /* SliceDivider (>= 1); rows with row % divider != offset are skipped */
  uint32_t seed;                        /* Seed (frameSeed, Raytracer.cs:660) — the Philox key */
  rtb_view view;                        /* View */
  rtb_environment environment;          /* Environment */
  uint32_t sample_count_range[2];       /* SampleCountRange (each <= 2^20) */
  int32_t trace_depth;                  /* TraceDepth (1..500, Raytracer.cs:90; accepted: 1..65535) */
  uint32_t sub_pixel_jitter;            /* SubPixelJitter (bool) */
  float sample_count_weight_extrema[2]; /* SampleCountWeightExtrema */
  /* Extension (not a SampleBatchJob field): restrict the batch to rows [row_begin, row_end).
   * Both 0 = all rows.  Used for row-tile sharding across GPUs; composes with the slice test. */
  int32_t row_begin;
  int32_t row_end;
} rtb_batch_params;

/* Diagnostics under FULL_DIAGNOSTICS (Raytracer.cs:54-64; the define is on in
 * ProjectSettings.asset:590).  ray_count matches the reference exactly (one per bounce-loop
 * iteration, SampleBatchJob.cs:203).  bounds_hit_count / candidate_count are what THIS
 * traversal executed (a pruned closest-hit walk visits fewer nodes than the reference's
 * collect-all walk, SampleBatchJob.cs:420-447). */
typedef struct rtb_diagnostics {
  float ray_count;
  float bounds_hit_count;
  float candidate_count;
  float sample_count_weight;
} rtb_diagnostics;

/* The accumulation buffers of one batch (SampleBatchJob.cs:41-51).  All eight point to
 * W*H elements.  `diagnostics` may be NULL.  in_* and out_* may not alias (the reference
 * ping-pongs two sets, Raytracer.cs:798-802). */
typedef struct rtb_batch_buffers {
  const float* in_color;                /* float4: xyz = sum of radiance, w = #successful samples */
  const float* in_sample_count_weight;  /* float */
  const float* in_normal;               /* float3 */
  const float* in_albedo;               /* float3 */
  float* out_color;                     /* float4 */
  float* out_sample_count_weight;       /* float */
  float* out_normal;                    /* float3 */
  float* out_albedo;                    /* float3 */
  rtb_diagnostics* out_diagnostics;     /* may be NULL */
} rtb_batch_buffers;

typedef struct rtb_ctx rtb_ctx;

/* ---- lifecycle ------------------------------------------------------------------------- */
RTB_API int rtb_abi_version(void);
/* Creates a context bound to CUDA device `device`.  Fails with RTB_ERR_CUDA+code when no
 * sm_100 device/driver is present: there is no CPU fallback. */
RTB_API int rtb_create(int device, rtb_ctx** out_ctx);
RTB_API int rtb_destroy(rtb_ctx* ctx);
/* Last error text of this context (or of the calling thread when ctx is NULL). */
RTB_API const char* rtb_last_error(const rtb_ctx* ctx);
typedef void (*rtb_log_fn)(int level, const char* message, void* user);
RTB_API int rtb_set_log_callback(rtb_ctx* ctx, rtb_log_fn fn, void* user);

/* ---- scene ----------------------------------------------------------------------------- */
/* Copies the flattened world to the device (replaces BvhRoot + the pointer graph behind it,
 * SampleBatchJob.cs:34).  The host may free its arrays on return. */
RTB_API int rtb_upload_scene(rtb_ctx* ctx,
                             const rtb_sphere* spheres, size_t sphere_count,
                             const rtb_material* materials, size_t material_count,
                             const rtb_bvh_node* nodes, size_t node_count);

/* The general form: leaves of `nodes` index into `entities`, each of which names a sphere or a triangle.
 * rtb_upload_scene(spheres, ...) is the same call with entities[i] = {SPHERE, i}. */
RTB_API int rtb_upload_world(rtb_ctx* ctx,
                             const rtb_entity* entities, size_t entity_count,
                             const rtb_sphere* spheres, size_t sphere_count,
                             const rtb_triangle* triangles, size_t triangle_count,
                             const rtb_material* materials, size_t material_count,
                             const rtb_bvh_node* nodes, size_t node_count);

/* rtb_upload_world plus placed entities: entities[i] of type RECT, BOX, or any type | RTB_ENTITY_PLACED index
 * `placed`.  Worlds with placed entities run the kernel flavour that carries entity transforms and the ray's
 * time; worlds without are exactly rtb_upload_world. */
RTB_API int rtb_upload_placed_world(rtb_ctx* ctx,
                                    const rtb_entity* entities, size_t entity_count,
                                    const rtb_sphere* spheres, size_t sphere_count,
                                    const rtb_triangle* triangles, size_t triangle_count,
                                    const rtb_placed_entity* placed, size_t placed_count,
                                    const rtb_material* materials, size_t material_count,
                                    const rtb_bvh_node* nodes, size_t node_count);

/* Image textures (TextureType.Image, Runtime/Texture.cs:80-89,128-137) — what the reference's host builds for mesh
 * materials whose Unity material carries a base-colour / emissive / metallic-gloss map (Raytracer.cs:1210-1267).
 * A material's four textures (Albedo, Emission, Glossiness, Metallic; Material.cs:22) are each either the constant in
 * rtb_material or an image whose texel scales that constant: colour = rgb / 255 * MainColor, scalar =
 * texel[channel] / 255 * MainColor[channel] (for a scalar the rtb_material value is the MainColor the host passes,
 * Raytracer.cs:1249,1258).  The texel is (int2)(TexCoords * ImageSize), row-major from `pixels`; the reference does
 * not wrap or clamp (it reads past the image for a coordinate of exactly 1): this plugin clamps to the image.
 * HitRecord.TexCoords are the triangle's interpolated vertex coordinates (HitTests.cs:147); every other entity
 * type has TexCoords = 0 (Entity.cs:107). */
typedef struct rtb_image {
  const uint8_t* pixels;        /* height rows of width texels of pixel_stride bytes */
  int32_t width, height;
  int32_t pixel_stride;         /* 3 = RGB24, 4 = RGBA32 */
  int32_t reserved;
} rtb_image;

typedef struct rtb_material_textures {  /* per material; image index or -1 = constant */
  int32_t albedo_image;
  int32_t emission_image;
  int32_t glossiness_image;
  int32_t metallic_image;
  int32_t glossiness_channel;   /* Texture.ScalarValueChannel (3 = the alpha of an RGBA32 map, Raytracer.cs:1260-1265) */
  int32_t metallic_channel;
  int32_t reserved[2];
} rtb_material_textures;        /* 32 bytes */

/* Attaches images to the world of the last rtb_upload_* call: material_count must equal that world's, triangle_uvs
 * (6 floats per triangle: the t1, t2, t3 of the Triangle ctor, Triangle.cs:14-29; NULL = all zero) its triangle
 * count.  Everything is copied.  A later rtb_upload_scene / _world / _placed_world drops the textures. */
RTB_API int rtb_upload_textures(rtb_ctx* ctx, const rtb_image* images, size_t image_count,
                                const rtb_material_textures* material_textures, size_t material_count,
                                const float* triangle_uvs, size_t triangle_count);

/* Environment.SkyCubemap (Runtime/Texture.cs:141-211; the host builds it from the scene's HDRI sky,
 * Raytracer.cs:663-665): six faces of R16G16B16A16_SFloat texels — the only format the reference accepts
 * (Texture.cs:155-163) — in CubemapFace order +X, -X, +Y, -Y, +Z, -Z, each face_height rows of face_width
 * texels of 4 halves.  Copied to the device; sampled (nearest texel, Cubemap.Sample) by batches whose
 * environment.sky_type is RTB_SKY_CUBEMAP.  NULL removes it. */
RTB_API int rtb_upload_sky_cubemap(rtb_ctx* ctx, const uint16_t* half_rgba, int face_width, int face_height);

/* How rtb_upload_scene would lay this world out on the device (no device needed; for tests and
 * tuning).  `leaf_spheres` as RTB_OPT_LEAF_SPHERES. */
typedef struct rtb_scene_layout {
  uint32_t inner_nodes;         /* device inner nodes (two child boxes each) */
  uint32_t leaves;              /* device leaves */
  uint32_t device_spheres;      /* sphere slots in depth-first leaf order (= spheres referenced by the BVH) */
  uint32_t max_leaf_spheres;
  uint32_t max_depth;           /* deepest root-to-leaf path of the device tree */
  uint32_t blob_bytes;          /* bytes staged to shared memory per CTA */
  uint32_t chain_boxes;         /* host boxes kept for exact re-testing of collapsed leaves */
  uint32_t collapsed;           /* 1 when some device leaf is a collapsed host subtree */
} rtb_scene_layout;
RTB_API int rtb_describe_scene(const rtb_sphere* spheres, size_t sphere_count,
                               const rtb_material* materials, size_t material_count,
                               const rtb_bvh_node* nodes, size_t node_count,
                               int leaf_spheres, rtb_scene_layout* out);

/* The tree the device walks under RTB_OPT_RETREE (no device needed; for tests and hosts that want to see it): a
 * surface-area-heuristic topology over the NON-EMPTY LEAVES of `nodes` — every leaf record (bounds, first_entity,
 * entity_count) is copied unchanged, inner bounds are unions of what lies below them, the root is node 0, nodes are in
 * depth-first order, no leaf lies deeper than 62 inner nodes.  Writes at most `capacity` nodes (2 * leaves - 1 <=
 * node_count are needed) and their count.  Returns RTB_OK with *out_count = 0 when the world does not qualify (a leaf box that is
 * flat, inverted or not finite, an inner box of `nodes` that does not contain its children's, a malformed tree, fewer than
 * two non-empty leaves): the host's topology is walked then.  Why the image cannot depend on the topology: csrc/retree.hpp. */
RTB_API int rtb_retree_bvh(const rtb_bvh_node* nodes, size_t node_count,
                           rtb_bvh_node* out_nodes, size_t capacity, size_t* out_count);

/* ---- the hot path ---------------------------------------------------------------------- */
/* Replaces `sampleBatchJob.Schedule(W*H, 1, dep).Complete()` (Raytracer.cs:730-736) with HOST
 * buffers: returns when the out_* arrays are fully written.  Pinned arrays (rtb_register_host_buffer /
 * cudaHostAlloc) are read and written in place by the megakernel over PCIe; pageable ones are staged
 * through device copies of the four inputs and the four outputs (+diagnostics).
 * `cancel` (may be NULL) is the CancellationToken (SampleBatchJob.cs:23,61; the host flips it through a raw
 * pointer, Raytracer.cs:189-192, then Complete()s, :512-515).  The batch is ONE kernel launch with or without a
 * token: the kernel polls a word in device memory owned by the context (a volatile load whenever a warp claims a
 * work tile; a CTA that saw it stops issuing samples), and the blocking call watches *cancel while it waits and
 * writes that word from a second stream when the token is set.  Once set, the call returns RTB_ERR_CANCELLED within a fraction of a millisecond
 * (the collect-all kernel of worlds with participating media: within one pixel's samples); outputs are then unspecified, as in the
 * reference (the host discards them).
 *
 * Accumulation range.  Per pixel and batch the sums of colour, normal, albedo and sampleCountWeight are kept in
 * signed 64-bit fixed point with 32 fraction bits (integer addition is associative: the image does not depend
 * on the order in which paths retire, on tiling or on the GPU count).  Consequences, which the reference's
 * float sums do not share: (1) a sample contributes round-to-nearest multiples of 2^-32 (components below 2^-33
 * contribute 0); (2) a successful sample with a non-finite component or one of magnitude >= 2^25 (3.3e7), or a
 * batch whose per-pixel sum reaches 2^30 (1.07e9) in magnitude, writes NaN to that pixel's out_color / out_normal / out_albedo /
 * out_sample_count_weight (CombineJob turns NaN into black, CombineJob.cs:50-53) — it never wraps silently. */
RTB_API int rtb_sample_batch(rtb_ctx* ctx, const rtb_batch_params* params,
                             const rtb_batch_buffers* host_buffers,
                             const volatile uint8_t* cancel);

/* Same batch on DEVICE-resident buffers, enqueued on `cuda_stream` (a cudaStream_t passed as
 * void*; NULL = the CUDA default stream) without synchronising: the caller owns the
 * buffers and the stream (used for multi-batch accumulation that never leaves HBM, and for
 * row-tile sharding where each rank renders into its slice of a frame).  The buffers may live on ANOTHER device
 * that this context's device can reach (cudaDeviceEnablePeerAccess, or a CUDA IPC mapping from rtb_ipc_open):
 * the kernel reads its rows' inputs and writes its rows' outputs over NVLink, so no gather follows it.
 * All batches of one context may be enqueued on different streams concurrently (each launch takes its own work
 * counter); instrumented batches (RTB_OPT_COUNTERS) must not overlap each other. */
RTB_API int rtb_sample_batch_device(rtb_ctx* ctx, const rtb_batch_params* params,
                                    const rtb_batch_buffers* device_buffers,
                                    void* cuda_stream);

/* Optional: pin (cudaHostRegister) the host's pooled accumulation arrays; the same 8 pointers
 * recur every batch (Raytracer.cs:279-303 pools).  When every array of a batch is pinned,
 * rtb_sample_batch runs the kernel directly on them (no staging copies, see RTB_OPT_HOST_ACCESS);
 * otherwise pinned arrays still copy at full PCIe rate. */
RTB_API int rtb_register_host_buffer(rtb_ctx* ctx, void* ptr, size_t bytes);
RTB_API int rtb_unregister_host_buffer(rtb_ctx* ctx, void* ptr);

/* ---- one host, N GPUs ------------------------------------------------------
 * The reference schedules ONE job over W*H pixels (Raytracer.cs:730); a host with several GPUs keeps that one call
 * site: an rtb_multi handle owns one context per device and renders one batch of the frame with all of them.  Every
 * pixel is independent and the Philox stream is keyed by the global pixel index, so device g simply takes the rows
 * [bounds[g], bounds[g + 1]) of the SAME buffers and the image is bit-identical for any device count.  There is no
 * gather: host arrays are written in place by every device over its own PCIe link (pinned) or staged per device
 * (pageable); device arrays live on one GPU and the others write their rows into them over NVLink peer access.
 * Row tiles are balanced inside the plugin: an instrumented probe batch (a few samples per pixel, split over the
 * devices) gives a per-row cost model whenever the world, the size or the view changes, and every batch's measured
 * per-device kernel times correct it (sky rows are cheap, glass is dear). */
#define RTB_MULTI_MAX_DEVICES 16
typedef struct rtb_multi rtb_multi;
/* `devices`: CUDA device ordinals.  (An ordinal may repeat: two contexts then share that GPU — of use on a one-GPU box
 * to exercise the tiling.) */
RTB_API int rtb_multi_create(const int* devices, int device_count, rtb_multi** out_multi);
RTB_API int rtb_multi_destroy(rtb_multi* m);
RTB_API int rtb_multi_device_count(const rtb_multi* m);
/* The context of device `index` of the handle (options, counters, rtb_last_kernel_ms); owned by the handle. */
RTB_API rtb_ctx* rtb_multi_context(rtb_multi* m, int index);
RTB_API const char* rtb_multi_last_error(const rtb_multi* m);
/* rtb_set_option on every device; RTB_OPT_BALANCE_TILES belongs to the handle itself. */
RTB_API int rtb_multi_set_option(rtb_multi* m, int option, int64_t value);
/* The rtb_upload_* calls, replicated to every device (the world is <= a few MB). */
RTB_API int rtb_multi_upload_scene(rtb_multi* m, const rtb_sphere* spheres, size_t sphere_count,
                                   const rtb_material* materials, size_t material_count,
                                   const rtb_bvh_node* nodes, size_t node_count);
RTB_API int rtb_multi_upload_placed_world(rtb_multi* m, const rtb_entity* entities, size_t entity_count,
                                          const rtb_sphere* spheres, size_t sphere_count,
                                          const rtb_triangle* triangles, size_t triangle_count,
                                          const rtb_placed_entity* placed, size_t placed_count,
                                          const rtb_material* materials, size_t material_count,
                                          const rtb_bvh_node* nodes, size_t node_count);
RTB_API int rtb_multi_upload_textures(rtb_multi* m, const rtb_image* images, size_t image_count,
                                      const rtb_material_textures* material_textures, size_t material_count,
                                      const float* triangle_uvs, size_t triangle_count);
RTB_API int rtb_multi_upload_sky_cubemap(rtb_multi* m, const uint16_t* half_rgba, int face_width, int face_height);
/* Pins a pooled host array once for every device of the handle (see rtb_register_host_buffer). */
RTB_API int rtb_multi_register_host_buffer(rtb_multi* m, void* ptr, size_t bytes);
RTB_API int rtb_multi_unregister_host_buffer(rtb_multi* m, void* ptr);
/* rtb_sample_batch over every device of the handle: blocking, HOST buffers, the same CancellationToken contract
 * (the token is relayed to every device's kernel).  params->row_begin/row_end restrict the batch as usual. */
RTB_API int rtb_multi_sample_batch(rtb_multi* m, const rtb_batch_params* params,
                                   const rtb_batch_buffers* host_buffers, const volatile uint8_t* cancel);
/* rtb_sample_batch_device over every device of the handle: the buffers live on device `owner_index` (index into the
 * handle's device list) and the batch is ordered on `owner_stream` of that device like any other work enqueued
 * there: it starts when the stream reaches it and the stream continues when every device's rows have landed.  The
 * other devices read and write the owner's memory over peer access (fails with RTB_ERR_UNSUPPORTED without a peer
 * path).  Does not synchronise. */
RTB_API int rtb_multi_sample_batch_device(rtb_multi* m, const rtb_batch_params* params,
                                          const rtb_batch_buffers* device_buffers, int owner_index, void* owner_stream);
/* Row bounds (device_count + 1 ints) and kernel milliseconds (device_count floats) of the last batch; either may be
 * NULL.  Kernel times of a device-buffer batch are available once the owner's stream has passed it. */
RTB_API int rtb_multi_get_tiles(rtb_multi* m, int* out_bounds, float* out_kernel_ms);
/* The balancer's partition, exposed for tests and for hosts that shard by themselves: bounds of `device_count` row
 * tiles of [row_begin, row_end) with near-equal sums of row_cost (indexed by absolute row; NULL = equal rows). */
RTB_API int rtb_balance_rows(const double* row_cost, int row_begin, int row_end, int device_count, int* out_bounds);

/* One process per GPU (torch.distributed-style hosts): device memory one process allocates and the other ranks map,
 * so that every rank's kernel writes its row tile straight into rank 0's frame over NVLink (the buffers of
 * rtb_sample_batch_device may be such a mapping).  Thin wrappers over cudaMalloc / cudaIpc*. */
typedef struct rtb_ipc_handle { unsigned char bytes[64]; } rtb_ipc_handle;
RTB_API int rtb_device_alloc(rtb_ctx* ctx, size_t bytes, void** out_device_ptr);     /* zero-filled */
RTB_API int rtb_device_free(rtb_ctx* ctx, void* device_ptr);
RTB_API int rtb_ipc_export(rtb_ctx* ctx, void* device_ptr, rtb_ipc_handle* out_handle);
RTB_API int rtb_ipc_open(rtb_ctx* ctx, const rtb_ipc_handle* handle, void** out_device_ptr);
RTB_API int rtb_ipc_close(rtb_ctx* ctx, void* device_ptr);

/* ---- adjacent jobs, device-side ("next" rows f1/f2 of SURVEY.md §8) --------------------- */
/* CombineJob (CombineJob.cs:29-71): rgb = color.xyz / (int)color.w with the interlace
 * look-around, NaN -> 0, albedo / max(n,1), normalizesafe(normal / max(n,1)).
 * Device pointers; out_* are float3 arrays; any out may be NULL. */
RTB_API int rtb_combine_device(rtb_ctx* ctx, int width, int height, int debug_mode, int ldr_albedo,
                               const float* color4, const float* normal3, const float* albedo3,
                               float* out_color3, float* out_normal3, float* out_albedo3,
                               void* cuda_stream);

/* FinalizeTexturesJob (FinalizeTexturesJob.cs:23-55): saturate(LinearToGamma(x)) * 255 -> RGBA32 for the colour,
 * normal * 0.5 + 0.5 and albedo images (float3 in, 4 bytes per pixel out).  Device pointers; any in/out pair may be NULL. */
RTB_API int rtb_finalize_device(rtb_ctx* ctx, int width, int height,
                                const float* color3, const float* normal3, const float* albedo3,
                                uint32_t* out_color_rgba, uint32_t* out_normal_rgba, uint32_t* out_albedo_rgba,
                                void* cuda_stream);

/* ReduceMetricsJob (ReduceMetricsJob.cs:22-45) on device buffers. */
typedef struct rtb_metrics {
  int64_t total_ray_count;
  int64_t total_samples;
  float sample_count_weight_min, sample_count_weight_max;
  int32_t sample_count_min, sample_count_max;
} rtb_metrics;
RTB_API int rtb_reduce_metrics_device(rtb_ctx* ctx, int width, int height,
                                      const rtb_diagnostics* diagnostics, const float* color4,
                                      const float* sample_count_weight, rtb_metrics* out_host,
                                      void* cuda_stream);

/* ---- measurement ----------------------------------------------------------------------- */
/* Work counters of the last rtb_sample_batch*_ call made with counters enabled
 * (rtb_set_option(ctx, RTB_OPT_COUNTERS, 1) selects the instrumented kernel build). */
typedef struct rtb_counters {
  uint64_t samples;             /* camera paths attempted */
  uint64_t rays;                /* bounce-loop iterations (== sum of ray_count) */
  uint64_t node_tests;          /* AABB slab tests executed */
  uint64_t sphere_tests;        /* ray-sphere quadratics executed */
  uint64_t shade_standard;      /* Standard scatter evaluations */
  uint64_t shade_dielectric;    /* Dielectric scatter evaluations */
  uint64_t sky_hits;            /* paths terminated by the sky */
  uint64_t failed_samples;      /* depth == TraceDepth (SampleBatchJob.cs:379-381) */
} rtb_counters;
RTB_API int rtb_get_counters(rtb_ctx* ctx, rtb_counters* out);

typedef enum rtb_option {
  RTB_OPT_COUNTERS = 1,         /* 0/1: run the instrumented kernel (slower) */
  RTB_OPT_KERNEL = 2,           /* 0 = auto, 1 = simple (thread per pixel), 2 = persistent megakernel */
  /* 3 was RTB_OPT_CANCEL_CHUNK_ROWS (ABI 1): the token is polled inside one launch now */
  RTB_OPT_LEAF_SPHERES = 4,     /* 1..15 (default 1): subtrees of the host's BVH holding at most this many spheres are
                                 * walked as one leaf on the device (results are identical for every value; takes effect
                                 * at the next rtb_upload_scene) */
  RTB_OPT_HOST_ACCESS = 6,      /* 1 (default): rtb_sample_batch lets the kernel read/write PINNED host arrays in place over PCIe
                                 * (rtb_register_host_buffer or cudaHostAlloc memory); 0: always stage through device copies */
  RTB_OPT_NOISE = 7,            /* 0 (default): Philox4x32-10 keyed (pixel, sample, bounce); 1: the reference's own white noise
                                 * (NoiseColor.White: one sequential Unity.Mathematics.Random xorshift32 stream per pixel per batch,
                                 * SampleBatchJob.cs:91) — validation only, needs RTB_OPT_KERNEL = 1 (a sequential stream cannot be
                                 * split over lanes) */
  RTB_OPT_MATH = 9,             /* 0 (default): the parity build — strict IEEE evaluation with FMAs only where written, every path
                                 * decision checkable bit for bit against the CPU restatement.  1: the fast build of the sphere
                                 * megakernel, compiled the way the reference's own [BurstCompile(FloatPrecision.Medium,
                                 * FloatMode.Fast)] (SampleBatchJob.cs:16) allows: FMA contraction, hardware sqrt / divide / sincos
                                 * approximations, one-FMA slab planes.  Same estimator and random numbers; images differ from the
                                 * parity build by a few flipped decisions per million paths (tools/fast_math_report.py).  Worlds
                                 * with triangles, placed entities, textures or media, and instrumented batches, keep the parity
                                 * kernels whatever this says. */
  RTB_OPT_RETREE = 10,          /* 1 (default): worlds of spheres and triangles without media are walked through a surface-area-heuristic
                                 * tree built over the LEAVES of the host's BVH (same leaf boxes, same entity ranges; rtb_retree_bvh)
                                 * instead of the host's topology, when every leaf box has min < max on every axis and every inner box
                                 * of the host's tree contains its children's.  An inner box of the reference's tree is the exact
                                 * union of its children's (BvhNodeData.cs:205-212) and the slab test is monotonic in the box, so a ray
                                 * that passes a leaf's box passes its whole chain: the reference's candidates are the entities of the
                                 * leaves whose own box is hit, whatever lies above them (csrc/retree.hpp) — same
                                 * image bit for bit, fewer boxes per ray (config 3: 125.0 -> 112.6 ms, mesh world 69.2 -> 61.6).
                                 * Entities at EXACTLY the same distance (triangles that share an edge) go to the one the reference's
                                 * candidate order puts first, whatever tree is walked.  2: also worlds with placed entities (Rect,
                                 * Box, rotated / moving entities): their kernel flavour leaves an exact tie to the entity visited
                                 * first, so 1 path in 1e8 may come out differently from the host's topology — hence opt-in.  Media
                                 * worlds always keep the host's topology.  0: walk the host's topology.  Takes effect at the next
                                 * upload; the traversal counters of instrumented batches count the walk that ran */
  RTB_OPT_BALANCE_TILES = 8,    /* rtb_multi only. 1 (default): cost-model + kernel-time balanced row tiles; 0: equal row counts */
  RTB_OPT_ALWAYS_WALK_CHAINS = 5 /* test knob, 0/1: re-test the host boxes a collapsed leaf skipped for EVERY accepted hit
                                 * instead of only when the hit geometry does not already prove them (same results, slower) */
} rtb_option;
RTB_API int rtb_set_option(rtb_ctx* ctx, int option, int64_t value);
/* 1 when the last rtb_sample_batch ran on the host arrays in place (all of them pinned), 0 when it staged copies. */
RTB_API int rtb_last_batch_in_place(rtb_ctx* ctx, int* out_in_place);
/* Milliseconds of the last kernel launch sequence on the context stream (CUDA events; the
 * same bracket as the two RecordTimeJobs, Raytracer.cs:729-738).  Valid after a synchronising call. */
RTB_API int rtb_last_kernel_ms(rtb_ctx* ctx, float* out_ms);
/* Roofline denominator for this path (MEASURED_PEAKS.json has no FP32-pipe figure):
 * runs an all-SM microbenchmark of independent FP32 FMA chains and returns the best of
 * `repeats` timings in TFLOP/s (FMA = 2 flop).  Not part of the reference's interface. */
RTB_API int rtb_measure_fp32_peak(rtb_ctx* ctx, int repeats, double* out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
