/*
 * rtb_host.h — C ABI of librtb_host.so: the HOST-side producers of the sample job's
 * inputs.  In the reference these live in the C# MonoBehaviour and its helpers and
 * "the host keeps them" (BASELINE.json north_star); there is no Unity host in this
 * repo, so the same steps are restated here in C++ (no CUDA, no oracle code) so that
 * the GPU plugin and the CPU oracle are fed byte-identical inputs.
 *
 * What each entry point mirrors (paths relative to
 * /root/reference/RaytracingInOneWeekend/Assets/Scripts):
 *   rtbh_random_*        Unity.Mathematics.Random (xorshift32; package not vendored)
 *   rtbh_scene_*         the legacy SceneData loader + random-group generator
 *                        (Unity/Raytracer.cs:1355-1506, commented out at HEAD) fed with
 *                        Assets/Scenes/Legacy/{Three Spheres,Final Scene} (Book 1).asset
 *   rtbh_build_bvh       Unity/BvhNodeData.cs:23-80,122-213,240-250 + Runtime/Jobs/BuildRuntimeBvhJob.cs:18-39
 *   rtbh_make_view       Runtime/View.cs:16-36 as called from Unity/Raytracer.cs:604-612
 *   rtbh_hit_world       Raytracer.HitWorld -> HitTests.Hit(this BvhNode) (Runtime/HitTests.cs:152-196), auto-focus
 *   rtbh_space_filling_series   Util/Tools.cs:101-124 (interlacing offsets, Raytracer.cs:650-661)
 */
#ifndef RTB_HOST_H
#define RTB_HOST_H

#include "rtb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Unity.Mathematics.Random ---------------------------------------------------------- */
typedef struct rtbh_random { uint32_t state; } rtbh_random;
RTB_API void rtbh_random_init(rtbh_random* r, uint32_t seed);       /* state = seed; NextState() */
RTB_API uint32_t rtbh_random_next_state(rtbh_random* r);            /* returns the OLD state */
RTB_API float rtbh_random_next_float(rtbh_random* r);               /* asfloat(0x3f800000 | s >> 9) - 1 */

/* ---- scenes ---------------------------------------------------------------------------- */
typedef struct rtbh_camera {
  float position[3];
  float target[3];
  float aperture;
  float vertical_fov;           /* degrees */
} rtbh_camera;

typedef struct rtbh_scene_info {
  rtbh_camera camera;
  rtb_environment environment;
  uint32_t sphere_count;
  uint32_t material_count;
  uint32_t lambertian_count, metal_count, dielectric_count;
  uint32_t tentative_draws;     /* dart-throwing iterations consumed */
} rtbh_scene_info;

typedef enum rtbh_scene_id {
  RTBH_SCENE_THREE_SPHERES = 0, /* Three Spheres (Book 1).asset */
  RTBH_SCENE_FINAL = 1,         /* Final Scene (Book 1).asset: 4 named + dart-thrown group, seed 700 */
  RTBH_SCENE_STRESS = 2         /* same generator, spread 110x110, draws until `target_count` accepted */
} rtbh_scene_id;

/* Fills spheres (scene order, NOT yet BVH order; material index i == sphere index i) and
 * materials.  Pass NULL arrays to query counts in `info`.  target_count is used by
 * RTBH_SCENE_STRESS only (number of accepted random spheres). */
RTB_API int rtbh_scene_generate(int scene_id, uint32_t seed, uint32_t target_count,
                                rtb_sphere* spheres, size_t sphere_capacity,
                                rtb_material* materials, size_t material_capacity,
                                rtbh_scene_info* info);

/* ---- BVH build + flatten ---------------------------------------------------------------- */
/* max_depth == 0 gives a single root leaf holding every sphere (the "linear hit list").
 * out_spheres receives the BVH-ordered entity copy (bvhEntities); node capacity needed is
 * at most 2*n-1 (>= 1).  Returns RTB_ERR_INVALID_ARGUMENT when a capacity is too small. */
RTB_API int rtbh_build_bvh(const rtb_sphere* spheres, size_t sphere_count, int max_depth,
                           rtb_sphere* out_spheres, size_t out_sphere_capacity,
                           rtb_bvh_node* out_nodes, size_t node_capacity, size_t* out_node_count);

/* The same builder over precomputed world bounds (6 floats per entity: min.xyz, max.xyz), for worlds
 * that mix entity types (CreateBvhBuildingEntitiesJob computes exactly these, BvhNodeData.cs:94-107).
 * out_order[i] = input index of the i-th entity of the BVH-ordered list (bvhEntities). */
RTB_API int rtbh_build_bvh_from_bounds(const float* bounds, size_t entity_count, int max_depth,
                                       uint32_t* out_order, size_t order_capacity,
                                       rtb_bvh_node* out_nodes, size_t node_capacity, size_t* out_node_count);
/* Entity bounds as the reference computes them: Sphere.Bounds through the (identity-rotation) rigid
 * transform (Sphere.cs:16-23, BvhNodeData.cs:41-78); Triangle.Bounds (Triangle.cs:38-49). */
RTB_API void rtbh_sphere_bounds(const rtb_sphere* sphere, float out_bounds[6]);
RTB_API void rtbh_triangle_bounds(const rtb_triangle* triangle, float out_bounds[6]);
/* A placed entity's world bounds (BvhBuildingEntity ctor, BvhNodeData.cs:28-80): the content's local box
 * (Sphere.cs:16-23, Rect.cs:17-19, Box.cs:17) through OriginTransform, swept over the motion when moving. */
RTB_API void rtbh_placed_bounds(const rtb_placed_entity* entity, float out_bounds[6]);
/* Triangle constructors (Triangle.cs:14-29): n1..n3 NULL = the face-normal form. */
RTB_API void rtbh_make_triangle(const float v1[3], const float v2[3], const float v3[3],
                                const float* n1, const float* n2, const float* n3,
                                uint32_t material, rtb_triangle* out);

/* Mesh ingestion (Runtime/Jobs/AddMeshRuntimeEntitiesJob.cs:30-96): one world-space triangle per index triple of a
 * Unity mesh (16-bit indices), the renderer's rigid transform and uniform scale (csum(lossyScale) / 3,
 * Raytracer.cs:1296) baked into the vertices, vertex normals rotated; normals == NULL selects the face-normal
 * ctor (MeshData.FaceNormals).  uvs (2 per vertex, TexCoord0) may be NULL; out_uvs (6 per triangle, for
 * rtb_upload_textures) may be NULL. */
RTB_API int rtbh_add_mesh(const float* vertices, const float* normals, const float* uvs, size_t vertex_count,
                          const uint16_t* indices, size_t index_count,
                          const float rotation[4], const float position[3], float scale, uint32_t material,
                          rtb_triangle* out_triangles, float* out_uvs, size_t triangle_capacity, size_t* out_triangle_count);

/* ---- camera ----------------------------------------------------------------------------- */
RTB_API void rtbh_make_view(const float origin[3], const float look_at[3], const float up[3],
                            float vertical_fov_degrees, float aspect, float aperture,
                            float focus_distance, rtb_view* out);
/* First hit along a ray through the whole BVH; returns 1 and the distance, or 0. */
RTB_API int rtbh_hit_world(const rtb_bvh_node* nodes, size_t node_count,
                           const rtb_sphere* spheres, size_t sphere_count,
                           const float origin[3], const float direction[3], float* out_distance);
/* ScheduleSample's camera block (Raytracer.cs:604-612) for a legacy-asset camera:
 * forward = normalize(target - position), up = (0,1,0), focus = first hit along forward
 * (else `fallback_focus`, the reference's initial 1). */
RTB_API void rtbh_view_from_camera(const rtbh_camera* camera, float aspect,
                                   const rtb_bvh_node* nodes, size_t node_count,
                                   const rtb_sphere* spheres, size_t sphere_count,
                                   float fallback_focus, rtb_view* out, float* out_focus_distance);

/* ---- interlacing ------------------------------------------------------------------------ */
RTB_API int rtbh_space_filling_series(int length, int32_t* out, size_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* RTB_HOST_H */
