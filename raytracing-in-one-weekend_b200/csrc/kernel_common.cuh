// kernel_common.cuh — device-side data layout, Philox RNG, TMA staging helpers and the
// per-bounce building blocks (camera ray, closest-hit traversal, Material.Scatter) shared
// by the kernels in sample_kernels.cuh.
//
// Reference paths are relative to /root/reference/RaytracingInOneWeekend/Assets/Scripts.
//
// Arithmetic contract (see include/rtb/umath.h): this translation unit is compiled with
// -fmad=false, so an FMA exists only where um::fma / um::dot / um::mad is written.  Every
// float expression below is written in the evaluation order of the CPU oracle so that the
// discrete decisions of a path (hit / miss, reflect / refract, which sphere is nearest)
// are bit-identical; the only tolerated differences are in sums and products whose order
// the kernels change on purpose (per-pixel accumulation, attenuation product).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rtb.h"
#include "rtb/umath.h"

namespace rtbk {

using um::f3;

// ---------------------------------------------------------------------------------------
// Scene blob: one contiguous, 16-byte aligned byte array in HBM, copied verbatim into
// shared memory by TMA bulk copies at CTA start (or read in place through the read-only
// path when it does not fit).
//
//   inner nodes  : n_inner * kNodeStride B   child boxes stored in the parent, 4 x float4 (+ padding up to the stride):
//                    q0 = (Lmin.x, Lmin.y, Lmax.x, Lmax.y)   x and y of every corner as an aligned pair, the four z's as two
//                    q1 = (Rmin.x, Rmin.y, Rmax.x, Rmax.y)   pairs: the operands of the packed slab arithmetic (aabb_range_pair)
//                    q2 = (Lmin.z, Lmax.z, Rmin.z, Rmax.z)
//                    q3 = (left_ref, right_ref, media flags, -) as int32
//                  ref >= 0: inner node, byte offset of its record (index * kNodeStride) — in the copy staged to shared memory
//                  the parity build turns these into addresses in the shared window once per launch (SceneView::rebias_inner_refs);
//                  ref < 0: leaf, ~ref = byte offset of its first slot in `spheres` | (count - 1)
//                  (count - 1 == 15: the count is leaf_count[first slot])
//                  Entities are named by the byte offset of their slot ("slot" below) everywhere in the kernels.
//   spheres      : n_spheres * 16 B  float4 (center.xyz, radius), depth-first leaf order
//   leaf_count   : n_spheres * 4 B   sphere count of the leaf that STARTS at this sphere
//   mat_index    : n_spheres * 4 B   material index of the sphere
//   triangles    : n_triangles * 80 B  5 x float4 per EntityType.Triangle entity (world space):
//                    T0 = (Data[0].xyz, Data[1].x)  T1 = (Data[1].yz, Data[2].xy)  T2 = (Data[2].z, n0.xyz)
//                    T3 = (n1.xyz, n2.x)            T4 = (n2.yz, -, -)          (Triangle.cs:10-11)
//                  a triangle entity's slot in `spheres` is (triangle index as bits, 0, 0, NaN): NaN radius = triangle
//   placed       : n_placed * 112 B  7 x float4 per entity with the reference's full Entity record (rtb_placed_entity:
//                  rotated or moving spheres, EntityType.Rect, EntityType.Box):
//                    P0 = OriginTransform.rot (x, y, z, w)      P1 = (OriginTransform.pos, type | moving << 8 as bits)
//                    P2 = InverseTransform.rot                   P3 = (InverseTransform.pos, -)      (static entities)
//                    P4 = (DestinationOffset, TimeRange.x)       P5 = (TimeRange.y, c0, c1, c2)      P6 = (c3, c4, c5, -)
//                  content c: Sphere (radius); Rect (From.xy, To.xy); Box (Extents.xyz, InverseExtents.xyz)
//                  a placed entity's slot in `spheres` is (placed index as bits, 1 as bits, 0, NaN)
// A device leaf is a subtree of the host's BVH with at most RTB_OPT_LEAF_SPHERES spheres
// (plugin.cu: Flattener); the host boxes it no longer walks are in chain_ref / chain_boxes (HBM).
// Materials (n_materials * 64 B DevMaterial) stay in HBM behind the read-only path: they are
// read once per bounce, not once per node.
// ---------------------------------------------------------------------------------------
struct DevMaterial {            // 64 bytes = 4 x float4; lives in HBM (read-only path), one read per bounce
  float albedo[3];
  uint32_t type;                // rtb_material_type
  float emission[3];
  float glossiness;
  float metallic;
  float ior;                    // Standard: lerp(PlasticIor, MetalIor, metallic) (Material.cs:86); Dielectric: IndexOfRefraction
  uint32_t perfect_specular;    // Material.IsPerfectSpecular (Material.cs:181-196)
  float roughness;              // Standard: pow(1 - glossiness, 2) (:82); Dielectric: 1 - glossiness (:123)
  // Per-material constants of Scatter, evaluated ON THE DEVICE at upload by the same functions the
  // per-hit code would call (derive_materials_kernel), so hoisting them changes no bits.
  float alpha;                  // Microfacet.RoughnessToAlpha(roughness) (Standard)
  float r0;                     // Schlick: ((1 - ior) / (1 + ior))^2
  float one_minus_r0;
  uint32_t textured;            // some texture of this material is an image (rtb_upload_textures): resolve_textures per hit
};
static_assert(sizeof(DevMaterial) == 64, "DevMaterial layout");

struct SceneDesc {
  const unsigned char* blob;    // device pointer
  uint32_t blob_bytes;          // multiple of 16
  uint32_t inner_off, sphere_off, leaf_count_off, mat_index_off, tri_off, placed_off;
  const DevMaterial* materials; // device pointer
  uint32_t n_inner, n_spheres, n_materials, n_triangles, n_placed;
  int32_t root_ref;             // as child refs; meaningful when has_root
  uint32_t has_root;            // 0: empty world (node_count == 0)
  float root_min[3], root_max[3];
  uint32_t max_depth;           // deepest root-to-leaf path (stack bound)
  // image textures (rtb_upload_textures); all in HBM, read only on hits of textured materials
  const unsigned char* tex_pixels;   // every image, back to back
  const int4* tex_images;            // per image: (byte offset, width, height, pixel stride)
  const int4* mat_textures;          // per material, 2 x int4: (albedo, emission, glossiness, metallic image or -1), (gloss channel, metallic channel, -, -)
  const float2* tri_uv;              // 3 per triangle (Triangle.TextureCoordinates columns) or nullptr
  const uint16_t* sky_faces;    // Environment.SkyCubemap texels (6 faces of RGBA halves) or nullptr
  int sky_w, sky_h;
  uint32_t has_volumes;         // some entity wears a ProbabilisticVolume: batches run the megakernel's media flavour (media.cuh) or, as the
                                // bit-exact validator / for the white-noise stream / with image textures, sample_volumes (volume_kernel.cuh)
  uint32_t has_big_leaves;      // some leaf holds 16 or more entities (count code 15: the count is in leaf_count)
  uint32_t has_chains;          // 1: some leaf is a collapsed subtree, accepted hits go through chain_guard; 2: always walk the chain (test knob)
  const uint32_t* chain_ref;    // per sphere: first chain box | box count << 24
  const float4* chain_boxes;    // 2 x float4 per box (min.xyz, max.xyz), tightest first
  uint32_t* status;             // device word of sticky kStatus* bits a kernel raises (read back by the synchronising calls)
  uint32_t skip_root_test;      // the tree is retree.hpp's: the root is an inner node whose box is the exact union of its children's, and a
                                // ray that misses a box misses every box inside it (retree.hpp) — the root's own slab test decides nothing
};
constexpr uint32_t kStatusHitListOverflow = 1u;   // sample_volumes: a ray met more entities than kMaxRayHits (volume_kernel.cuh)

// Bytes between inner-node records.  The walk reads a node as four LDS.128 at [ref + 16 k]; with 64-byte records the
// k-th quad of EVERY node starts in one of two 16-byte bank groups (64 B = 16 banks), so divergent lanes pile up on
// 8 of the 32 banks (ncu: 31 % of the shared wavefronts are conflict replays).  An odd multiple of 16 bytes (80, 112)
// spreads them over all eight bank groups — measured on config 3: no faster (80: 151.3 vs 151.4 ms, 129.9 vs 130.2;
// 112: 155.2; the LSU pipe is at 20 % and the replays hide behind the ALU work), so the records stay packed.  Refs are
// byte offsets, the walk does not depend on the stride.
#ifndef RTB_NODE_STRIDE
#define RTB_NODE_STRIDE 64
#endif
constexpr uint32_t kNodeStride = RTB_NODE_STRIDE;
static_assert(kNodeStride >= 64 && kNodeStride % 16 == 0, "node records are 4 x float4, 16-byte aligned");

constexpr int kDefaultCollapse = 1;   // measured on B200 (profiles/README.md): single-sphere leaves are fastest for the staged-in-smem walk
// Geometry of the guard that stands in for the skipped boxes (see chain_guard).
constexpr float kChainShrink = 1.0e-3f;   // upload checks: every skipped box contains its sphere shrunk by this

constexpr int kStackMax = 64;   // traversal stack entries (BVH depth bound; upload rejects deeper trees)
constexpr int kTilePixelsMax = 16;
constexpr uint32_t kPhiloxKey1 = 0x52544232u;   // "RTB2"
constexpr uint32_t kBounceCamera = 0xFFFFFFFFu;

// Kernel arguments (one __grid_constant__ struct).
struct BatchArgs {
  rtb_batch_params p;
  rtb_batch_buffers b;          // DEVICE pointers
  SceneDesc scene;
  int width, height;
  int first_row, row_step, n_rows;   // active rows: first_row + j * row_step, j < n_rows
  uint32_t n_active_pixels;          // n_rows * width
  // Work tiles, claimed in order from tile_counter.  Three phases of shrinking tile size (guided
  // self-scheduling): phase k covers active pixels [phase_pixel[k], phase_pixel[k+1]) in tiles of
  // phase_size[k] pixels (<= kTilePixelsMax), starting at tile id phase_tile[k].  Big tiles amortise the
  // per-tile drain; the small ones at the end keep the kernel's tail (warps finishing their last tile
  // at different times) short.
  uint32_t phase_tile[4], phase_pixel[4];
  int phase_size[3];
  uint32_t n_tiles;
  uint32_t refill_min;               // megakernel: free lanes a warp waits for before it refills them (plugin.cu: choose_tiles)
  uint32_t* tile_counter;            // global work counter of THIS launch (zeroed before it; plugin.cu keeps a ring of them)
  // CancellationToken (SampleBatchJob.cs:61 polls it per pixel): a word in DEVICE memory owned by the context (never
  // NULL).  The blocking call watches the caller's token while the kernel runs and, when it is set, writes this batch's
  // epoch into the word from a second stream; the kernel reads the word (volatile, an L2 hit) whenever a warp claims a
  // tile and stops issuing work once it equals cancel_epoch.  (Epochs: a late write of an earlier batch never matches.
  // Device memory, not mapped host memory: thousands of warps polling one host word serialise on the PCIe round trip —
  // measured 2.5x on the whole kernel.)
  const uint32_t* cancel_flag;
  uint32_t cancel_epoch;
  unsigned long long* counters;      // rtb_counters as 8 x u64, or nullptr
};

__device__ __forceinline__ void tile_range(const BatchArgs& a, uint32_t t, uint32_t* base, int* n) {
  const int k = t >= a.phase_tile[2] ? 2 : (t >= a.phase_tile[1] ? 1 : 0);
  const uint32_t b = a.phase_pixel[k] + (t - a.phase_tile[k]) * (uint32_t)a.phase_size[k];
  *base = b;
  *n = (int)min((uint32_t)a.phase_size[k], a.phase_pixel[k + 1] - b);
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter = (pixel, sample, bounce, block),
// key = (Seed, "RTB2").  Slots: see DESIGN.md "Random-draw slots".
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// RandomSource.NextFloat (RandomSource.cs:130-151 -> Unity.Mathematics.Random.NextFloat)
__device__ __forceinline__ float u2f(uint32_t u) { return um::asfloat(0x3f800000u | (u >> 9)) - 1.0f; }

// ---------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------
// Scene accessors.  SMEM = true: `base` points into shared memory (plain loads become LDS);
// SMEM = false: `base` is the blob in HBM and loads go through the read-only path.
// ---------------------------------------------------------------------------------------
template <bool SMEM>
struct SceneView {
  // SMEM: `s` is the blob's address in the shared window, kept as a 32-bit register so that every access is a
  // plain LDS [reg + imm] (a generic pointer makes the compiler rebuild the window base around each load);
  // else `g` is the blob in HBM.
  const unsigned char* g;
  uint32_t s;
  uint32_t inner_off, sphere_off, leaf_count_off, mat_index_off, tri_off, placed_off;

  __device__ __forceinline__ void bind(const unsigned char* base, const SceneDesc& d) {
    g = base;
    s = SMEM ? smem_u32(base) : 0u;
#ifndef RTB_NO_OPAQUE_SMEM_BASE
    // opaque to the optimiser: otherwise it rebuilds the window address (S2UR CgaCtaId, UMOV, ULEA) around every group
    // of loads instead of keeping it in a register (3 of the walk's 65 instructions per node; measured 148.4 -> 144.5 ms)
    asm volatile("mov.u32 %0, %0;" : "+r"(s));
#endif
    inner_off = d.inner_off; sphere_off = d.sphere_off; leaf_count_off = d.leaf_count_off; mat_index_off = d.mat_index_off;
    tri_off = d.tri_off;
    placed_off = d.placed_off;
  }
  __device__ __forceinline__ float4 triangle(uint32_t index, int k) const { return ld4(tri_off + index * 80u + (uint32_t)k * 16u); }
  __device__ __forceinline__ float4 placed(uint32_t index, int k) const { return ld4(placed_off + index * 112u + (uint32_t)k * 16u); }
  __device__ __forceinline__ float4 ld4(uint32_t off) const {
    if (SMEM) {
      float4 v;
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s + off));
      return v;
    }
    return __ldg(reinterpret_cast<const float4*>(g + off));
  }
  __device__ __forceinline__ uint32_t ld1(uint32_t off) const {
    if (SMEM) {
      uint32_t v;
      asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(s + off));
      return v;
    }
    return __ldg(reinterpret_cast<const uint32_t*>(g + off));
  }
  // `ref` = an inner child ref: the byte offset of the node record (the inner section starts the blob).  RTB_ABS_REFS: in the
  // staged copy the inner refs are ADDRESSES in the shared window (rebias_inner_refs adds the blob's address once per launch),
  // so a node load is LDS [ref + imm] without the add of the window base per visit.
  // Measured (bit-identical): config 3 114.0 -> 112.6 ms, mesh world 69.2 -> 68.4, Cornell box 114.5 -> 114.4, fog 352.2 -> 353.6;
  // the fast build 97.2 -> 97.8, so it keeps offsets.
#ifndef RTB_ABS_REFS
#ifdef RTB_FAST_MATH
#define RTB_ABS_REFS 0
#else
#define RTB_ABS_REFS 1
#endif
#endif
  __device__ __forceinline__ float4 node(int ref, int k) const {
    if (SMEM && RTB_ABS_REFS) {
      float4 v;
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)ref + (uint32_t)k * 16u));
      return v;
    }
    return ld4((uint32_t)ref + (uint32_t)k * 16u);
  }
  __device__ __forceinline__ int root(const SceneDesc& d) const {
    return (SMEM && RTB_ABS_REFS && d.root_ref >= 0) ? d.root_ref + (int)s : d.root_ref;
  }
  // once per launch, by every thread of the CTA, after the staged copy has landed and before anyone walks (caller syncs)
  __device__ __forceinline__ void rebias_inner_refs(uint32_t n_inner, uint32_t tid, uint32_t n_threads) const {
    if (!(SMEM && RTB_ABS_REFS)) return;
    for (uint32_t i = tid; i < n_inner; i += n_threads) {
      const uint32_t at = s + inner_off + i * kNodeStride + 48u;
      int l, r;
      asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(l), "=r"(r) : "r"(at));
      if (l >= 0) l += (int)s;
      if (r >= 0) r += (int)s;
      asm volatile("st.shared.v2.s32 [%0], {%1, %2};" :: "r"(at), "r"(l), "r"(r) : "memory");
    }
  }
  // `slot` = byte offset of an entity's 16-byte slot in the blob
  __device__ __forceinline__ float4 sphere(int slot) const { return ld4((uint32_t)slot); }
  __device__ __forceinline__ uint32_t leaf_count(int slot) const { return ld1(leaf_count_off + (((uint32_t)slot - sphere_off) >> 2)); }
  __device__ __forceinline__ uint32_t material_of(int slot) const { return ld1(mat_index_off + (((uint32_t)slot - sphere_off) >> 2)); }
};

struct WorkCounters {           // per-thread tallies of the instrumented build
  uint32_t node_tests = 0, sphere_tests = 0, shade_standard = 0, shade_dielectric = 0;
};

__device__ __forceinline__ f3 v3(const float* p) { return um::mk(p[0], p[1], p[2]); }

__device__ __forceinline__ bool cancel_requested(const uint32_t* flag, uint32_t epoch);
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ bool cancel_requested(const uint32_t* flag, uint32_t epoch) { return ld_volatile_u32(flag) == epoch; }

// HitTests.Hit(this AxisAlignedBoundingBox) (HitTests.cs:9-21): returns the decision and tMin.
// fminf/fmaxf agree with math.min/max ("isnan(y) || x < y ? x : y") on every input except
// the sign of a zero result, which no comparison below can observe.
__device__ __forceinline__ bool aabb_hit(f3 mn, f3 mx, f3 o, f3 inv, float* t_enter) {
  f3 t0 = (mn - o) * inv;
  f3 t1 = (mx - o) * inv;
  float tmin = fmaxf(0.0f, fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z)));
  float tmax = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
  *t_enter = tmin;
  return tmin < tmax;
}

// Kernel flavours (template int FLAVOR): what the walk has to handle beyond single-sphere leaves of spheres.
constexpr int kFlavorSpheres = 0;        // spheres only, no collapsed leaves, every leaf < 16 spheres (the fast build of tree worlds)
constexpr int kFlavorChains = 1;         // + collapsed leaves: accepted hits go through chain_guard; + leaves of 16 or more spheres
                                         // (a linear hit list is ONE such leaf).  Its own build because the tree walk sits on a
                                         // register / instruction-fetch cliff: the big-leaf code merely being present cost the
                                         // tree worlds 16 % (130 -> 151 ms on config 3, measured; profiles/README.md)
constexpr int kFlavorGeneral = 2;        // + EntityType.Triangle entities
constexpr int kFlavorPlaced = 3;         // + placed entities: rotation, motion (Ray.Time), EntityType.Rect, EntityType.Box
constexpr int kFlavorPlacedTextured = 4; // the same with image textures (the general flavour always has them; here the
                                         // texture code costs the untextured walk 14 %, so it is its own instantiation)
constexpr int kFlavorMedia = 5;          // every entity kind (no image textures) + MaterialType.ProbabilisticVolume: the walk is replaced
                                         // by the media hit search of media.cuh (pruned_hits + the reference's volume bookkeeping)
__host__ __device__ constexpr bool flavor_has_textures(int flavor) { return flavor == kFlavorGeneral || flavor == kFlavorPlacedTextured; }

// Ray.Time (View.cs:47; kept by every scattered ray, Material.cs:95-159): one draw per camera path, slot 4 of the
// camera ray's draws = word 0 of Philox block 1.  Only moving entities read it, so the megakernel re-derives it from
// the path's counters when one is tested instead of carrying it; the per-pixel kernel (whose white-noise mode has to
// draw it in stream order anyway) passes the value.
struct RayClock {
  uint32_t pixel, sample, seed;
  float value;
  bool known;
  __device__ __forceinline__ float time() const {
    if (known) return value;
    return u2f(philox4x32_10(pixel, sample, kBounceCamera, 1u, seed, kPhiloxKey1).x);
  }
};

// The host boxes between a collapsed device leaf and sphere `idx`, applied exactly as the reference
// would (FindHitCandidates reaches a sphere only through a chain of hit boxes, SampleBatchJob.cs:420-447).
__device__ __noinline__ bool chain_boxes_hit(const SceneDesc& sd, int slot, f3 o, f3 inv) {
  const uint32_t ref = __ldg(sd.chain_ref + (((uint32_t)slot - sd.sphere_off) >> 4));
  const float4* b = sd.chain_boxes + 2 * (size_t)(ref & 0xffffffu);
  for (uint32_t k = ref >> 24; k > 0; k--, b += 2) {
    const float4 mn = __ldg(b), mx = __ldg(b + 1);
    float t;
    if (!aabb_hit(um::mk(mn.x, mn.y, mn.z), um::mk(mx.x, mx.y, mx.z), o, inv, &t)) return false;
  }
  return true;
}

// When does an accepted sphere hit PROVE that every skipped box test passes?  Every skipped box
// contains the sphere shrunk by kChainShrink (checked at upload).  If the ray passes the centre at a
// distance <= (1 - 2 kChainShrink) |r|, the chord midpoint (t_m = -b/a >= 0) lies at least
// kChainShrink |r| inside every face of every such box, so each slab interval contains
// t_m -/+ kChainShrink |r| / |d|; the slab arithmetic (one subtraction, one reciprocal, one product:
// relative error < 2^-22 per bound) cannot close an interval that wide while t_m |d| < 2^10 |r|.
// In terms of the rounded values the sphere test already has: disc / a = r^2 - dist^2, computed with an
// absolute error below 2^-20 a (|oc|^2 + r^2).  Hits that fail the guard (grazing, very distant, or
// leaving a sphere the ray started in) take chain_boxes_hit.
__device__ __forceinline__ bool chain_guard(float a, float b, float oc2, float r2, float disc) {
  return b <= 0.0f && disc >= a * (8.0e-3f * r2 + 2.0e-6f * (oc2 + r2)) && b * b <= 1.0e6f * (a * r2);
}

// The same slab arithmetic returning both ends: the box is hit iff *t_enter < *t_exit.
__device__ __forceinline__ void aabb_range(f3 mn, f3 mx, f3 o, f3 inv, float* t_enter, float* t_exit) {
  f3 t0 = (mn - o) * inv;
  f3 t1 = (mx - o) * inv;
  *t_enter = fmaxf(0.0f, fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z)));
  *t_exit = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
}

// Inner-node records (64 bytes; plugin.cu: Flattener::ref_of) keep the two child boxes so that x and y of every corner form
// an aligned pair and the four z's two pairs:
//   q0 = (L.min.x, L.min.y, L.max.x, L.max.y)   q1 = (R.min.x, R.min.y, R.max.x, R.max.y)
//   q2 = (L.min.z, L.max.z, R.min.z, R.max.z)   q3 = (left ref, right ref, media flags, -)
__device__ __forceinline__ f3 node_lmin(float4 q0, float4 q1, float4 q2) { return um::mk(q0.x, q0.y, q2.x); }
__device__ __forceinline__ f3 node_lmax(float4 q0, float4 q1, float4 q2) { return um::mk(q0.z, q0.w, q2.y); }
__device__ __forceinline__ f3 node_rmin(float4 q0, float4 q1, float4 q2) { return um::mk(q1.x, q1.y, q2.z); }
__device__ __forceinline__ f3 node_rmax(float4 q0, float4 q1, float4 q2) { return um::mk(q1.z, q1.w, q2.w); }

// Blackwell's packed FP32 instructions (add / mul / fma .f32x2: SASS FADD2 / FMUL2 / FFMA2) do two IEEE round-to-nearest
// operations on a 64-bit register pair in ONE issue slot — and issue slots, not the FMA pipe, are what the walk runs out
// of (ncu: 87 % issue, FMA pipe 31 %).  Each half is the scalar instruction's result bit for bit, so the slab test of a
// visit — 12 subtractions and 12 products — is 6 + 6 packed instructions with the same 24 roundings.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct RayPairs { f32x2 oxy, ozz, ixy, izz; };     // (o.x, o.y), (o.z, o.z), (inv.x, inv.y), (inv.z, inv.z)
// aabb_range of both child boxes of a node record: the same (mn - o) * inv, (mx - o) * inv and min / max as two calls of it
__device__ __forceinline__ void aabb_range_pair(float4 q0, float4 q1, float4 q2, const RayPairs& rp,
                                                float* tl, float* xl, float* tr, float* xr) {
  float l0x, l0y, l1x, l1y, r0x, r0y, r1x, r1y, l0z, l1z, r0z, r1z;
  unpack2(mul2(sub2(pack2(q0.x, q0.y), rp.oxy), rp.ixy), l0x, l0y);
  unpack2(mul2(sub2(pack2(q0.z, q0.w), rp.oxy), rp.ixy), l1x, l1y);
  unpack2(mul2(sub2(pack2(q1.x, q1.y), rp.oxy), rp.ixy), r0x, r0y);
  unpack2(mul2(sub2(pack2(q1.z, q1.w), rp.oxy), rp.ixy), r1x, r1y);
  unpack2(mul2(sub2(pack2(q2.x, q2.y), rp.ozz), rp.izz), l0z, l1z);
  unpack2(mul2(sub2(pack2(q2.z, q2.w), rp.ozz), rp.izz), r0z, r1z);
  *tl = fmaxf(0.0f, fmaxf(fmaxf(fminf(l0x, l1x), fminf(l0y, l1y)), fminf(l0z, l1z)));
  *xl = fminf(fminf(fmaxf(l0x, l1x), fmaxf(l0y, l1y)), fmaxf(l0z, l1z));
  *tr = fmaxf(0.0f, fmaxf(fmaxf(fminf(r0x, r1x), fminf(r0y, r1y)), fminf(r0z, r1z)));
  *xr = fminf(fminf(fmaxf(r0x, r1x), fmaxf(r0y, r1y)), fmaxf(r0z, r1z));
}

// The fast build's version (see aabb_range_fast): t = mn * inv + (-(o * inv)), one packed FMA per pair of planes; rp.oxy / rp.ozz
// hold -(o * inv).
__device__ __forceinline__ void aabb_range_pair_fast(float4 q0, float4 q1, float4 q2, const RayPairs& rp,
                                                     float* tl, float* xl, float* tr, float* xr) {
  float l0x, l0y, l1x, l1y, r0x, r0y, r1x, r1y, l0z, l1z, r0z, r1z;
  unpack2(fma2(pack2(q0.x, q0.y), rp.ixy, rp.oxy), l0x, l0y);
  unpack2(fma2(pack2(q0.z, q0.w), rp.ixy, rp.oxy), l1x, l1y);
  unpack2(fma2(pack2(q1.x, q1.y), rp.ixy, rp.oxy), r0x, r0y);
  unpack2(fma2(pack2(q1.z, q1.w), rp.ixy, rp.oxy), r1x, r1y);
  unpack2(fma2(pack2(q2.x, q2.y), rp.izz, rp.ozz), l0z, l1z);
  unpack2(fma2(pack2(q2.z, q2.w), rp.izz, rp.ozz), r0z, r1z);
  *tl = fmaxf(0.0f, fmaxf(fmaxf(fminf(l0x, l1x), fminf(l0y, l1y)), fminf(l0z, l1z)));
  *xl = fminf(fminf(fmaxf(l0x, l1x), fmaxf(l0y, l1y)), fmaxf(l0z, l1z));
  *tr = fmaxf(0.0f, fmaxf(fmaxf(fminf(r0x, r1x), fminf(r0y, r1y)), fminf(r0z, r1z)));
  *xr = fminf(fminf(fmaxf(r0x, r1x), fmaxf(r0y, r1y)), fmaxf(r0z, r1z));
}

#ifdef RTB_FAST_MATH
// The fast build's slab test: t = mn * inv - o * inv with o * inv hoisted out of the walk — one FMA per plane instead of a
// subtraction and a product (12 instead of 24 FP32 instructions per two-box visit).  Not the parity build's roundings.
__device__ __forceinline__ void aabb_range_fast(f3 mn, f3 mx, f3 inv, f3 noi, float* t_enter, float* t_exit) {
  f3 t0 = um::mad(mn, inv, noi);
  f3 t1 = um::mad(mx, inv, noi);
  *t_enter = fmaxf(0.0f, fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z)));
  *t_exit = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
}
#endif

// HitTests.Hit(this Sphere) (HitTests.cs:23-60) behind Entity.HitInternal (Entity.cs:74-103)
// for a static, unrotated entity: entity-space origin = o + (-center), direction unchanged.
// `a` = dot(d, d) is hoisted out by the caller.  Updates (best_t, best_idx) when this sphere
// is hit nearer than best_t: the same record FindHits' sort would put first
// (SampleBatchJob.cs:450-475) — a root is accepted iff 0 < t < +inf there, and the
// second root is never nearer than the first, so clipping at best_t changes nothing.
//
// DEFER (big leaves of the lean sphere builds — a linear hit list is one): the division by a = dot(d, d) is taken out of
// the leaf's loop.  Every ray of the job has |d| = 1 up to rounding; when |a - 1| <= 1e-4 (`a_ok`, decided once per walk)
// and nothing was hit before the leaf, its candidates are compared by their NUMERATORS x = -b -/+ sqrt(disc) and the
// leaf's single winner is divided once after the loop (measured on the linear list of config 2: 57.7 -> 52.1 ms; in
// single-sphere leaves of a tree the same trick LOST 16 %, so those keep dividing per candidate):
//   * x -> fl(x / a) is monotonic, so the smallest numerator has the smallest distance (numerators that differ and
//     round to the same distance are a tie the reference's unstable sort does not define either);
//   * fl(x / a) > 0 iff x > 0 (no float underflows when divided by a number in [0.9999, 1.0001]);
//   * the prune limit best * kPruneMargin stays beyond the hit: fl(x / a) <= x * 1.00011 < x * kPruneMargin.
// A ray with any other |d|, or one that already holds a hit, divides every candidate as before.
// Equal distances (two triangles that share an edge report the SAME float distance for a third of the rays that land on it,
// tools/tie_probe.py): the reference's candidate order and sort decide (FindHitCandidates pops the right child first, FindHits
// pops candidates from the end and sorts: SampleBatchJob.cs:420-475; the oracle restates the sort as a stable one).  Of two
// entities at the same distance the one in the LOWER slot wins across leaves and the HIGHER slot inside one leaf (slots are in
// the reference's depth-first leaf order: plugin.cu; media.cuh: visited_later).  A leaf's entities are tested in ascending
// order, so with `leaf_first` = the first slot of the leaf under test the newcomer wins a tie iff the holder's slot is not
// below this leaf.  The triangle flavour applies it — meshes are where ties happen, and there the two extra instructions of an
// accepted hit cost nothing (mesh world 69.2 -> 68.2 ms) — and its winner then depends neither on the walk's order nor on its
// tree; so does the per-pixel validation kernel (sample_simple) for every world.  kNoTies (plain comparison, the first entity
// visited keeps a tie): the megakernel's lean sphere builds, whose worlds' spheres do not intersect, and its placed-entity
// flavours, which are instruction-fetch bound (Cornell box 114.5 -> 118.4 ms with the rule).
constexpr int kNoTies = 0x7fffffff;
__device__ __forceinline__ bool nearer(float t, float best_t, int best_idx, int leaf_first) {
  return t < best_t || (leaf_first != kNoTies && t == best_t && best_idx >= leaf_first);
}

template <bool CHAINS, bool DEFER>
__device__ __forceinline__ void sphere_hit(const SceneDesc& sd, float4 s, int idx, f3 o, f3 d, f3 inv, float a,
                                           float& best_t, int& best_idx, const int leaf_first = kNoTies) {
  f3 oc = o + um::mk(-s.x, -s.y, -s.z);
  float b = um::dot(oc, d);
  float oc2 = um::dot(oc, oc);
  float r2 = s.w * s.w;
  float c = oc2 - r2;
  float disc = um::fma(b, b, -(a * c));
  // HitTests.cs:33-55 takes the first of t1 = (-b - sqrt(disc)) / a, t2 = (-b + sqrt(disc)) / a that lies in (0, best_t).
  // Cases decided without evaluating a root, with the SAME outcome as evaluating both (a > 0; rounding is monotonic,
  // and sqrt(fl(b*b)) == |b|, so c > 0 gives disc <= fl(b*b) and sqrt(disc) <= |b|):
  //   c > 0, b >= 0 (outside, moving away):  -b - sq <= 0 and -b + sq <= 0: neither root is > 0            -> no hit
  //   c <= 0 (inside or on the sphere):      sq >= |b| so t1 <= 0: only t2 can be accepted
  //   c > 0, b < 0:                          0 <= t1 <= t2: if t1 > 0 it alone decides (t1 >= best_t implies t2 >= best_t)
  if (disc > 0.0f && !(c > 0.0f && b >= 0.0f)) {
    const float sq = um::sqrt(disc);
    float t = c > 0.0f ? -b - sq : -b + sq;
    if (DEFER) {
      if (c > 0.0f && !(t > 0.0f)) t = -b + sq;                  // t1 rounded to 0: the reference moves on to t2
    } else {
      t = um::div(t, a);
      if (c > 0.0f && !(t > 0.0f)) t = um::div(-b + sq, a);
    }
    if (nearer(t, best_t, best_idx, leaf_first) && t > 0.0f) {
      if (CHAINS && sd.has_chains && (sd.has_chains == 2u || !chain_guard(a, b, oc2, r2, disc)) && !chain_boxes_hit(sd, idx, o, inv)) return;
      best_t = t;
      best_idx = idx;
    }
  }
}

// HitTests.Hit(this Triangle) (HitTests.cs:113-150): Moeller-Trumbore in the reference's operation order, both
// faces.  Returns u, v and the distance when the reference would report a hit for tMin = 0, tMax = +inf
// (FindHits, SampleBatchJob.cs:457): distance < 0 is the only rejected range.
template <bool SMEM>
__device__ __forceinline__ bool triangle_uvt(const SceneView<SMEM>& sv, uint32_t tri, f3 o, f3 d, float* u, float* v, float* t) {
  const float4 a = sv.triangle(tri, 0), b = sv.triangle(tri, 1), c = sv.triangle(tri, 2);
  const f3 data0 = um::mk(a.x, a.y, a.z), data1 = um::mk(a.w, b.x, b.y), data2 = um::mk(b.z, b.w, c.x);
  const f3 pvec = um::cross(d, data0);
  const float det = um::dot(data1, pvec);
  if (det == 0) return false;
  const float inv_det = um::div(1.0f, det);
  const f3 tvec = o - data2;
  *u = um::dot(tvec, pvec) * inv_det;
  if (*u < 0 || *u > 1) return false;
  const f3 qvec = um::cross(tvec, data1);
  *v = um::dot(d, qvec) * inv_det;
  if (*v < 0 || *u + *v > 1) return false;
  *t = um::dot(data0, qvec) * inv_det;
  return !(*t < 0.0f);
}
template <bool SMEM>
__device__ __forceinline__ void triangle_hit(const SceneView<SMEM>& sv, uint32_t tri, int idx, f3 o, f3 d, float& best_t, int& best_idx,
                                             const int leaf_first) {
  float u, v, t;
  if (triangle_uvt(sv, tri, o, d, &u, &v, &t) && nearer(t, best_t, best_idx, leaf_first)) {
    best_t = t;
    best_idx = idx;
  }
}
// math.sign
__device__ __forceinline__ float sign_of(float x) { return (x > 0 ? 1.0f : 0.0f) - (x < 0 ? 1.0f : 0.0f); }

// Entity.HitInternal + HitContent (Entity.cs:74-122) for a placed entity with tMin = 0, tMax = +inf (FindHits,
// SampleBatchJob.cs:457): the transform at the ray's time, the ray in entity space, then the Sphere / Rect / Box test
// (HitTests.cs:23-111) in the reference's operation order.  Returns the distance and the entity-space normal rotated
// back by the transform (not yet normalised).
// WITH_EXIT (media.cuh): also the hit FindHits injects behind a convex medium's entry (SampleBatchJob.cs:462-469) — Entity.Hit
// again with tMin = distance + 0.001 — from the same transform: for a sphere the same two roots decide it, a Box repeats its
// test from the offset origin (HitTests.cs:84), a Rect has none (not a convex hull).  *t_exit < 0: no such hit.
template <bool SMEM, bool WITH_EXIT>
__device__ __forceinline__ bool placed_test_core(const SceneView<SMEM>& sv, uint32_t pidx, f3 o, f3 d, const RayClock& clk,
                                                 float* t_out, f3* n_out, float tmin, float* t_exit, f3* n_exit) {
  const float4 p0 = sv.placed(pidx, 0), p1 = sv.placed(pidx, 1);
  const uint32_t flags = __float_as_uint(p1.w);
  um::quat rot;
  rot.x = p0.x; rot.y = p0.y; rot.z = p0.z; rot.w = p0.w;
  um::rigid inv;
  if ((flags >> 8) == 0u) {                     // static: InverseTransform from the Entity ctor (Entity.cs:51-52)
    const float4 p2 = sv.placed(pidx, 2), p3 = sv.placed(pidx, 3);
    inv.rot.x = p2.x; inv.rot.y = p2.y; inv.rot.z = p2.z; inv.rot.w = p2.w;
    inv.pos = um::mk(p3.x, p3.y, p3.z);
  } else {                                      // Entity.TransformAtTime (Entity.cs:124-127), inverted per ray (:87-88)
    const float4 p4 = sv.placed(pidx, 4);
    const float tr_y = sv.placed(pidx, 5).x;
    const float k = um::clamp(um::unlerp(p4.w, tr_y, clk.time()), 0.0f, 1.0f);
    um::rigid at;
    at.rot = rot;
    at.pos = um::mk(p1.x, p1.y, p1.z) + um::mk(p4.x, p4.y, p4.z) * k;
    inv = um::inverse(at);
  }
  const f3 eo = um::transform(inv, o);
  const f3 ed = um::rotate(inv.rot, d);
  const float4 p5 = sv.placed(pidx, 5);
  const uint32_t type = flags & 0xffu;
  float t;
  f3 n;
  if (WITH_EXIT) *t_exit = -1.0f;
  if (type == RTB_ENTITY_SPHERE) {              // HitTests.cs:23-60
    const float radius = p5.y;
    const float a = um::dot(ed, ed), b = um::dot(eo, ed), c = um::dot(eo, eo) - radius * radius;
    const float disc = um::fma(b, b, -(a * c));
    if (!(disc > 0)) return false;
    const float sq = um::sqrt(disc);
    t = um::div(-b - sq, a);
    bool first_root = true;
    if (!(t < um::INF && t > tmin)) {
      t = um::div(-b + sq, a);
      first_root = false;
      if (!(t < um::INF && t > tmin)) return false;
    }
    n = um::mad(ed, t, eo) / radius;
    if (WITH_EXIT && first_root) {
      // the second call finds the first root again (not beyond t + 0.001), then the second one
      const float t2 = um::div(-b + sq, a);
      if (t2 < um::INF && t2 > t + 0.001f) {
        *t_exit = t2;
        *n_exit = um::rotate(rot, um::mad(ed, t2, eo) / radius);
      }
    }
  } else if (type == RTB_ENTITY_RECT) {         // HitTests.cs:62-78
    if (ed.z >= 0) return false;
    t = um::div(-eo.z, ed.z);
    if (t < tmin || t > um::INF) return false;
    const float x = eo.x + t * ed.x, y = eo.y + t * ed.y;
    const float to_y = sv.placed(pidx, 6).x;
    if (x < p5.y || y < p5.z || x > p5.w || y > to_y) return false;
    n = um::mk(0.0f, 0.0f, 1.0f);
  } else {                                      // Box, HitTests.cs:80-111
    const float4 p6 = sv.placed(pidx, 6);
    const f3 ext = um::mk(p5.y, p5.z, p5.w), inv_ext = um::mk(p6.x, p6.y, p6.z);
    const f3 sgn = um::mk(-sign_of(ed.x), -sign_of(ed.y), -sign_of(ed.z));
    float from = tmin;
#pragma unroll 1
    for (int pass = 0; pass < (WITH_EXIT ? 2 : 1); pass++) {
      const f3 bo = eo + ed * from;             // "offset origin by tMin"
      const f3 ao = um::mk(um::abs(bo.x), um::abs(bo.y), um::abs(bo.z)) * inv_ext;
      const float winding = um::cmax(ao) < 1 ? -1.0f : 1.0f;
      const f3 num = ext * winding * sgn - bo;
      const f3 dp = um::mk(um::div(num.x, ed.x), um::div(num.y, ed.y), um::div(num.z, ed.z));
      const bool tx = dp.x >= 0 && um::abs(bo.y + ed.y * dp.x) < ext.y && um::abs(bo.z + ed.z * dp.x) < ext.z;
      const bool ty = dp.y >= 0 && um::abs(bo.z + ed.z * dp.y) < ext.z && um::abs(bo.x + ed.x * dp.y) < ext.x;
      const bool tz = dp.z >= 0 && um::abs(bo.x + ed.x * dp.z) < ext.x && um::abs(bo.y + ed.y * dp.z) < ext.y;
      const f3 bn = tx ? um::mk(sgn.x, 0.0f, 0.0f) : ty ? um::mk(0.0f, sgn.y, 0.0f) : um::mk(0.0f, 0.0f, tz ? sgn.z : 0.0f);
      const bool nzx = bn.x != 0, nzy = bn.y != 0, nzz = bn.z != 0;
      bool ok = nzx || nzy || nzz;
      float bt = nzx ? dp.x : nzy ? dp.y : dp.z;
      bt += from;
      if (bt > um::INF) ok = false;
      if (pass == 0) {
        if (!ok) return false;
        t = bt;
        n = bn;
        from = bt + 0.001f;
      } else if (ok) {
        *t_exit = bt;
        *n_exit = um::rotate(rot, bn);
      }
    }
  }
  *t_out = t;
  *n_out = um::rotate(rot, n);
  return true;
}
template <bool SMEM>
__device__ __noinline__ bool placed_test(const SceneView<SMEM>& sv, uint32_t pidx, f3 o, f3 d, const RayClock& clk,
                                         float* t_out, f3* n_out, float tmin = 0.0f) {
  return placed_test_core<SMEM, false>(sv, pidx, o, d, clk, t_out, n_out, tmin, nullptr, nullptr);
}
template <bool SMEM>
__device__ __forceinline__ void placed_hit(const SceneView<SMEM>& sv, uint32_t pidx, int idx, f3 o, f3 d, const RayClock& clk,
                                           float& best_t, int& best_idx, const int leaf_first) {
  float t;
  f3 n;
  if (placed_test(sv, pidx, o, d, clk, &t, &n) && nearer(t, best_t, best_idx, leaf_first)) {
    best_t = t;
    best_idx = idx;
  }
}

// HitRecord.Normal of the entity in slot `prim` hit at distance t (Entity.cs:57-72): sphere (HitTests.cs:41-45),
// triangle (interpolated vertex normals, HitTests.cs:144-147) or placed entity (the hit is re-derived, bit for bit,
// from the same ray).
template <bool SMEM, int FLAVOR>
__device__ __forceinline__ f3 hit_normal(const SceneView<SMEM>& sv, float4 prim, f3 o, f3 d, float t, const RayClock& clk) {
  constexpr bool TRIS = FLAVOR >= kFlavorGeneral;
  if (FLAVOR >= kFlavorPlaced && prim.w != prim.w && __float_as_uint(prim.y) != 0u) {
    float tt = 0;
    f3 n = um::mk(0.0f);
    placed_test(sv, __float_as_uint(prim.x), o, d, clk, &tt, &n);
    return um::normalize(n);
  }
  if (TRIS && prim.w != prim.w) {
    const uint32_t tri = __float_as_uint(prim.x);
    float u = 0, v = 0, tt = 0;
    triangle_uvt(sv, tri, o, d, &u, &v, &tt);
    const float4 c = sv.triangle(tri, 2), e = sv.triangle(tri, 3), f = sv.triangle(tri, 4);
    const f3 bary = um::mk(1 - u - v, u, v);
    return um::normalize(um::mul_cols(um::mk(c.y, c.z, c.w), um::mk(e.x, e.y, e.z), um::mk(e.w, f.x, f.y), bary));
  }
  const f3 oc = o + um::mk(-prim.x, -prim.y, -prim.z);
  return um::normalize(um::mad(d, t, oc) / prim.w);
}

// Closest hit over the whole world.  Replaces FindHitCandidates + FindHits
// (SampleBatchJob.cs:403-475): the reference collects every entity of every leaf whose box
// chain is hit, intersects all, sorts and takes index 0.  This walk applies the SAME box
// test to the same boxes and the same sphere test, but visits the nearer child first and
// skips a box whose entry distance lies beyond the best hit so far (with a safety margin
// far larger than the rounding error of either distance), so it returns the same nearest
// record while testing a fraction of the nodes.
// The lanes of a warp must meet at the top of EVERY trip of the walk loop.  Whether they do is the compiler's choice of loop
// structure — the same source has compiled into one loop (box visit executed with 20.3 of 32 lanes on config 3) and into a
// loop nest in which lanes iterate "visits that need no pop" on their own (13.8 lanes, +17 % kernel time), flipped by
// unrelated edits elsewhere in the kernel (profiles/README.md, round 2).  A convergent operation at the loop top pins the
// one-loop form: 4 instructions per trip (VOTE, R2UR, BRA.DIV, NOP).  Measured alternatives: the same operation at the
// loop's latch (132.8 vs 127.2 ms: the pop predicate then lives in a register across it), an uniform-exit loop on
// __any_sync (137.4), a branch-free pop that leaves the loop a single latch (139.8), empty inline asm / __activemask()
// alone (the nest comes back: 159.8).
#ifndef RTB_WALK_SYNC
#define RTB_WALK_SYNC 1
#endif
__device__ __forceinline__ void walk_converge() {
#if RTB_WALK_SYNC == 1
  __syncwarp(__activemask());
#endif
}

#ifndef RTB_PACKED_SLABS
#define RTB_PACKED_SLABS 1
#endif
constexpr float kPruneMargin = 1.0005f;
constexpr int kTraversalDone = (int)0x80000000;   // stack sentinel: neither an inner index (>= 0) nor ~first

#ifndef RTB_DEFER_DIV
#define RTB_DEFER_DIV 1   // lean sphere builds: a big leaf divides its winner once instead of every candidate (see sphere_hit)
#endif
// A leaf of 16 or more spheres in the lean builds (a linear hit list is one such leaf holding the world), out of line so
// that the tree walk's own code stays small (the megakernel is instruction-fetch sensitive: every few hundred instructions
// inlined into the walk cost milliseconds).  Deferred division (see sphere_hit) when the ray's |d|^2 is 1 to 1e-4 and nothing
// was hit before the leaf: candidates are compared by their numerators and the leaf's winner is divided once.
template <bool SMEM, bool CHAINS>
__device__ __noinline__ float2 big_leaf_hit(SceneView<SMEM> sv, const SceneDesc& sd, int first, int count, f3 o, f3 d, f3 inv, float a,
                                            float best_t, int best_idx) {
  if (RTB_DEFER_DIV && best_idx < 0 && um::abs(a - 1.0f) <= 1.0e-4f) {
#pragma unroll 4
    for (int i = 0; i < count; i++) {
      const int slot = first + 16 * i;
      sphere_hit<CHAINS, true>(sd, sv.sphere(slot), slot, o, d, inv, a, best_t, best_idx);
    }
    if (best_idx >= 0) {
      best_t = um::div(best_t, a);       // the winner's numerator -> its distance (HitTests.cs:33,45)
      if (!(best_t < um::INF)) { best_idx = -1; best_t = um::INF; }   // "t < tMax" with tMax = +inf (SampleBatchJob.cs:457)
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < count; i++) {
      const int slot = first + 16 * i;
      sphere_hit<CHAINS, false>(sd, sv.sphere(slot), slot, o, d, inv, a, best_t, best_idx);
    }
  }
  return make_float2(best_t, __int_as_float(best_idx));
}

template <bool SMEM, bool COUNTERS, int FLAVOR, bool TIES = (FLAVOR == kFlavorGeneral)>
__device__ __forceinline__ void closest_hit(const SceneView<SMEM>& sv, const SceneDesc& sd, f3 o, f3 d,
                                            float& best_t, int& best_idx, WorkCounters& wc, const RayClock& clk) {
  best_t = um::INF;
  best_idx = -1;
  if (!sd.has_root) return;
  // SampleBatchJob.cs:409-412
  f3 inv = um::rcp(d);
  inv = um::mk(um::isnan(inv.x) ? um::INF : inv.x, um::isnan(inv.y) ? um::INF : inv.y, um::isnan(inv.z) ? um::INF : inv.z);
  const float a = um::dot(d, d);

  // On a re-built tree the root's own slab test decides nothing (sd.skip_root_test).  Skipping it pays in the fast build only
  // (98.8 -> 97.2 ms on config 3; parity build 114.0 -> 114.3, 10 k spheres 35.7 -> 35.5: within noise, so it keeps the test).
#ifndef RTB_ROOT_SKIP
#define RTB_ROOT_SKIP 0
#endif
  if (!(RTB_ROOT_SKIP && sd.skip_root_test)) {
    float t_enter;
    if (COUNTERS) wc.node_tests++;
    if (!aabb_hit(v3(sd.root_min), v3(sd.root_max), o, inv, &t_enter)) return;
  }

  int stack[kStackMax];
  stack[0] = kTraversalDone;
  int cur = sv.root(sd);
  auto test_prim = [&](int slot, int leaf_first) {
    const float4 prim = sv.sphere(slot);
    const int ties = TIES ? leaf_first : kNoTies;
    if (FLAVOR >= kFlavorPlaced && prim.w != prim.w && __float_as_uint(prim.y) != 0u) placed_hit(sv, __float_as_uint(prim.x), slot, o, d, clk, best_t, best_idx, ties);
    else if (FLAVOR >= kFlavorGeneral && prim.w != prim.w) triangle_hit(sv, __float_as_uint(prim.x), slot, o, d, best_t, best_idx, ties);
    else sphere_hit<(FLAVOR >= kFlavorChains), false>(sd, prim, slot, o, d, inv, a, best_t, best_idx, ties);
  };
  auto test_leaf = [&](int ref) {
    const uint32_t code = (uint32_t)~ref;
    const int first = (int)(code & ~15u);
    if ((code & 15u) == 0u) {            // the common leaf: one entity (BvhNodeData.cs:155 splits down to n <= 1)
      test_prim(first, first);
      if (COUNTERS) wc.sphere_tests++;
      return;
    }
    int count = (int)(code & 15u) + 1;
    if (FLAVOR != kFlavorSpheres && count == 16) {   // a big leaf (upload sends worlds that have one to the other flavours)
      count = (int)sv.leaf_count(first);
      if (FLAVOR < kFlavorGeneral) {
        if (COUNTERS) wc.sphere_tests += count;
        const float2 r = big_leaf_hit<SMEM, (FLAVOR >= kFlavorChains)>(sv, sd, first, count, o, d, inv, a, best_t, best_idx);
        best_t = r.x;
        best_idx = __float_as_int(r.y);
        return;
      }
    }
#pragma unroll 1
    for (int i = 0; i < count; i++) test_prim(first + 16 * i, first);
    if (COUNTERS) wc.sphere_tests += count;
  };
  if (FLAVOR < kFlavorGeneral) {
  // The lean builds.  One trip = [box visit if the lane holds an inner node] -> [leaf test if it now holds a leaf] ->
  // [pop if it needs one]: a lane that descends into a leaf tests it in the same trip, and every lane passes the pop
  // once per trip (measured 131.4 -> 129.1 ms on config 3; the general flavour is faster with the split trips below).
  {
    int* top = stack + 1;
#if RTB_PACKED_SLABS && !defined(RTB_FAST_MATH)
    const RayPairs rp{pack2(o.x, o.y), pack2(o.z, o.z), pack2(inv.x, inv.y), pack2(inv.z, inv.z)};
#endif
#ifdef RTB_FAST_MATH
    // -(o * inv).  A direction component of exactly 0 gives inv = inf and inf - inf = NaN planes, which fminf / fmaxf drop:
    // that axis then constrains nothing (the box test errs on the side of visiting), the entity tests decide
    const f3 noi = um::mk(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
    const RayPairs rpf{pack2(noi.x, noi.y), pack2(noi.z, noi.z), pack2(inv.x, inv.y), pack2(inv.z, inv.z)};
#endif
    for (;;) {
      walk_converge();
      bool need_pop = false;
      if (cur >= 0) {
        const float4 q0 = sv.node(cur, 0), q1 = sv.node(cur, 1), q2 = sv.node(cur, 2), q3 = sv.node(cur, 3);
        float tl, tr, xl, xr;
#ifdef RTB_FAST_MATH
        aabb_range_pair_fast(q0, q1, q2, rpf, &tl, &xl, &tr, &xr);
#elif RTB_PACKED_SLABS
        aabb_range_pair(q0, q1, q2, rp, &tl, &xl, &tr, &xr);
#else
        aabb_range(node_lmin(q0, q1, q2), node_lmax(q0, q1, q2), o, inv, &tl, &xl);
        aabb_range(node_rmin(q0, q1, q2), node_rmax(q0, q1, q2), o, inv, &tr, &xr);
#endif
        const float limit = best_t * kPruneMargin;
        const bool hl = tl < fminf(xl, limit);
        const bool hr = tr < fminf(xr, limit);
        if (COUNTERS) wc.node_tests += 2;
        const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
        if (hl && hr) {
          const bool left_first = tl <= tr;
          *top++ = left_first ? right : left;
          cur = left_first ? left : right;
        } else if (hl || hr) {
          cur = hl ? left : right;
        } else {
          need_pop = true;
        }
      }
      if (!need_pop && cur < 0) {
        test_leaf(cur);
        need_pop = true;
      }
      if (need_pop) {
        cur = *--top;
        if (cur == kTraversalDone) break;
      }
    }
  }
    return;
  }
  // The general and placed flavours: one node (inner OR leaf) per trip (measured faster for them than the lean builds' fused
  // trips: 72.8 vs 78.8 ms on the mesh world).  Walking in rounds (collect 2 / 4 / 8 candidate entities, then intersect them with
  // the warp converged — what media.cuh's gather_hits does, where it pays) was measured here too: Cornell box 132.8 -> 139.4 /
  // 136.0 / 136.3 ms, as a linear list 99.3 -> 115.4 / 114.1 / 118.4, mesh world 69.4 -> 68.0 / 66.2 / 67.4 — not kept.
  int* top = stack + 1;
  const RayPairs rp{pack2(o.x, o.y), pack2(o.z, o.z), pack2(inv.x, inv.y), pack2(inv.z, inv.z)};
  for (;;) {
    walk_converge();
    if (cur >= 0) {
      const float4 q0 = sv.node(cur, 0), q1 = sv.node(cur, 1), q2 = sv.node(cur, 2), q3 = sv.node(cur, 3);
      // "box hit (t_enter < t_exit) and not beyond the best hit (t_enter < limit)" as ONE comparison per child
      // against min(t_exit, limit) (t_exit is never NaN: fminf/fmaxf drop NaN operands)
      float tl, tr, xl, xr;
      if (FLAVOR == kFlavorGeneral) {     // packed slabs: mesh world 69.7 -> 68.9 ms; the placed flavours lose with them (Cornell box 114.5 -> 116.1)
        aabb_range_pair(q0, q1, q2, rp, &tl, &xl, &tr, &xr);
      } else {
        aabb_range(node_lmin(q0, q1, q2), node_lmax(q0, q1, q2), o, inv, &tl, &xl);
        aabb_range(node_rmin(q0, q1, q2), node_rmax(q0, q1, q2), o, inv, &tr, &xr);
      }
      const float limit = best_t * kPruneMargin;
      const bool hl = tl < fminf(xl, limit);
      const bool hr = tr < fminf(xr, limit);
      if (COUNTERS) wc.node_tests += 2;
      const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
      if (hl && hr) {
        const bool left_first = tl <= tr;
        *top++ = left_first ? right : left;
        cur = left_first ? left : right;
        continue;
      }
      if (hl) { cur = left; continue; }
      if (hr) { cur = right; continue; }
    } else {
      test_leaf(cur);
    }
    cur = *--top;
    if (cur == kTraversalDone) break;
  }
}

// ---------------------------------------------------------------------------------------
// Sampling transforms (RandomSource.cs) and basis helpers (Util/Tools.cs:19-37)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ f3 tangent_to_world(f3 v, f3 normal) {
  float s = normal.z >= 0 ? 1.0f : -1.0f;
  float a = um::div(-1.0f, s + normal.z);
  float b = normal.x * normal.y * a;
  f3 tangent = um::mk(1 + s * normal.x * normal.x * a, s * b, -s * normal.x);
  f3 bitangent = um::mk(b, s + normal.y * normal.y * a, -normal.y);
  return um::normalize(um::mul_cols(tangent, normal, bitangent, v));
}
// The three sampling transforms take the angle as u * 2 * PI (RandomSource.cs:70), u * PI * 2
// (:125) and u * (2 * PI) (:46): doubling is exact, so all three are the same float and one
// sincos serves whichever transform a lane needs.
__device__ __forceinline__ void unit_angle_sincos(float u, float* s, float* c) { um::sincos(u * 2 * um::PI, s, c); }
// RandomSource.OnCosineWeightedHemisphere (RandomSource.cs:63-89); (s, c) = sincos(uy * 2 * PI)
__device__ __forceinline__ f3 cosine_hemisphere(f3 normal, float ux, float s, float c) {
  float radius = um::sqrt(ux);
  f3 tangent_space = um::mk(radius * c, um::sqrt(1 - ux), radius * s);
  return tangent_to_world(tangent_space, normal);
}
// RandomSource.NextFloat3Direction (RandomSource.cs:113-128); (s, c) = sincos(ry * PI * 2)
__device__ __forceinline__ f3 random_direction(float rx, float s, float c) {
  float z = rx * 2.0f - 1.0f;
  float r = um::sqrt(um::max(1.0f - z * z, 0.0f));
  return um::mk(c * r, s * r, z);
}

// Material.cs:212-217, split at the per-material constant
__device__ __forceinline__ float schlick_r0(float ior) {
  float r0 = um::div(1 - ior, 1 + ior);
  r0 *= r0;
  return r0;
}
__device__ __forceinline__ float schlick(float cosine, float r0, float one_minus_r0) {
  return r0 + one_minus_r0 * um::pow5(1 - cosine);
}
// Microfacet.cs:71-80
__device__ __forceinline__ float roughness_to_alpha(float roughness) {
  roughness = um::max(roughness, 1e-3f);
  float x = um::log(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
// Microfacet.cs:9-12,53-69; alpha = RoughnessToAlpha(roughness) hoisted to the material record
__device__ __forceinline__ float smith_masking_shadowing(f3 w, f3 normal, float alpha) {
  float cos_theta = um::dot(normal, w);
  float sq_cos = cos_theta * cos_theta;
  float sq_sin = um::max(0.0f, 1 - sq_cos);
  float sin_theta = um::sqrt(sq_sin);
  float tan_theta = um::div(sin_theta, cos_theta);
  float abs_tan = um::abs(tan_theta);
  float lambda;
  if (um::isinf(abs_tan)) {
    lambda = 0;
  } else {
    float a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
    lambda = um::div(-1 + um::sqrt(1 + a2t2), 2.0f);
  }
  return um::div(1.0f, 1 + lambda);
}

// ---------------------------------------------------------------------------------------
// One path vertex
// ---------------------------------------------------------------------------------------
struct PathRay {
  f3 o, d;
};

// View.GetRay (View.cs:38-48) with the uv of SampleBatchJob.cs:134.  The ray-time draw
// (View.cs:47, slot 4) is not generated here: only moving entities read Ray.Time (RayClock).
// Split in two so that a kernel can draw the Philox block and evaluate the sincos for camera rays and for bounces in
// ONE converged step (sample_megakernel): camera_ray_finish consumes the block r = Philox(pixel, sample, CAMERA, 0) and
// (s, c) = sincos(u2f(r.z) * 2 PI) — RandomSource.InUnitDisk's theta = u * (2 PI - 0) + 0 is the same float.
__device__ __forceinline__ PathRay camera_ray_finish(const rtb_batch_params& p, int cx, int cy, uint4 r, float s, float c) {
  const rtb_view& v = p.view;
  const bool lens = v.lens_radius != 0;
  float jx = 0.5f, jy = 0.5f, rdx = 0, rdy = 0;
  if (p.sub_pixel_jitter) { jx = u2f(r.x); jy = u2f(r.y); }
  if (lens) {  // RandomSource.InUnitDisk (RandomSource.cs:40-61)
    float radius = um::sqrt(u2f(r.w));
    rdx = v.lens_radius * (radius * c);
    rdy = v.lens_radius * (radius * s);
  }
  float nx = um::div((float)cx + jx, p.size[0]);
  float ny = um::div((float)cy + jy, p.size[1]);
  f3 offset = v3(v.right) * rdx + v3(v.up) * rdy;
  PathRay ray;
  ray.d = um::normalize(v3(v.lower_left_corner) - offset + nx * v3(v.horizontal) + ny * v3(v.vertical));
  ray.o = v3(v.origin) + offset;
  return ray;
}
__device__ __forceinline__ PathRay camera_ray(const rtb_batch_params& p, int cx, int cy, uint32_t pixel, uint32_t sample) {
  const bool lens = p.view.lens_radius != 0;
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  float s = 0, c = 1;
  if (p.sub_pixel_jitter || lens) {
    r = philox4x32_10(pixel, sample, kBounceCamera, 0, p.seed, kPhiloxKey1);
    if (lens) unit_angle_sincos(u2f(r.z), &s, &c);
  }
  return camera_ray_finish(p, cx, cy, r, s, c);
}

struct ScatterResult {
  f3 dir;           // scattered direction (not re-normalised, Material.cs:145,153)
  f3 reflectance;   // attenuation
  float random_events;
};

// Material.Scatter (Material.cs:67-173) for constant textures (Texture.cs:50-59,101-108).
// m0 = (albedo.xyz, type), m1 = (emission.xyz, glossiness), m2 = (metallic, ior, perfect_specular, roughness),
// m3 = (alpha, r0, 1 - r0, -).
//
// Layout for SIMT efficiency: a warp almost always holds Lambertian, metal and glass hits at once, so
// the pieces every material needs are executed ONCE, converged, with per-lane operands — one Philox
// block (its index is the only per-material difference), one sincos, one cosine-hemisphere sample —
// and only the short material-specific tails diverge.
struct ScatterPlan {   // what a bounce needs from the shared draw step: which Philox block, and whether sincos(u2f(r.y) * 2 PI)
  bool standard, pure_diffuse, need_angle;
  uint32_t block;
};
__device__ __forceinline__ ScatterPlan scatter_plan(float4 m0, float4 m1, float4 m2) {
  ScatterPlan pl;
  pl.standard = __float_as_uint(m0.w) == RTB_MATERIAL_STANDARD;   // else Dielectric (upload rejects the rest)
  const float glossiness = m1.w, metallic = m2.x, roughness = m2.w;
  // With glossiness == 0 the reflection chance saturate(fresnel * 0 * G) is exactly 0 and with
  // metallic == 0 the rough normal has no other consumer: the first draw pair (slots 0,1) is skipped
  // and the diffuse draw (block 1) is the only one — same result as evaluating Material.cs:83-89.
  pl.pure_diffuse = pl.standard && glossiness == 0.0f && metallic == 0.0f;
  pl.block = pl.pure_diffuse ? 1u : 0u;
  // Dielectric with roughness 0: normalize(N + 0 * randomDirection) == normalize(N); the direction is not evaluated.
  pl.need_angle = pl.standard ? (pl.pure_diffuse || roughness > 0) : (roughness > 0);
  return pl;
}
__device__ __forceinline__ ScatterResult scatter_finish(float4 m0, float4 m1, float4 m2, float4 m3, f3 D, f3 N, ScatterPlan pl,
                                                        uint4 r, float sn, float cs,
                                                        uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t seed, float ev0 = 0.0f);
// ev0: RandomSource.RandomEvents when Scatter starts (non-zero only after Material.ProbabilisticHit calls of the same bounce
// iteration; the reference keeps adding to the same float, so the sum has to start from it to round identically)
__device__ __forceinline__ ScatterResult scatter(float4 m0, float4 m1, float4 m2, float4 m3, f3 D, f3 N,
                                                 uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t seed, float ev0 = 0.0f) {
  const ScatterPlan pl = scatter_plan(m0, m1, m2);
  const uint4 r = philox4x32_10(pixel, sample, bounce, pl.block, seed, kPhiloxKey1);
  float sn = 0, cs = 1;
  if (pl.need_angle) unit_angle_sincos(u2f(r.y), &sn, &cs);
  return scatter_finish(m0, m1, m2, m3, D, N, pl, r, sn, cs, pixel, sample, bounce, seed, ev0);
}
__device__ __forceinline__ ScatterResult scatter_finish(float4 m0, float4 m1, float4 m2, float4 m3, f3 D, f3 N, ScatterPlan pl,
                                                        uint4 r, float sn, float cs,
                                                        uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t seed, float ev0) {
  ScatterResult out;
  out.reflectance = um::mk(m0.x, m0.y, m0.z);
  const bool standard = pl.standard, pure_diffuse = pl.pure_diffuse, need_angle = pl.need_angle;
  const float glossiness = m1.w, metallic = m2.x, roughness = m2.w;
  const float ux = u2f(r.x);
  f3 hemi = N;
  if (standard && need_angle) hemi = cosine_hemisphere(N, ux, sn, cs);

  if (standard) {
    float chance = 0.0f;
    f3 rough_normal = N;
    if (!pure_diffuse) {
      if (roughness > 0) rough_normal = um::normalize(um::lerp(N, hemi, roughness));
      const float incident_cosine = -um::dot(D, rough_normal);
      const float fresnel = schlick(incident_cosine, m3.y, m3.z);
      const float g = smith_masking_shadowing(D, N, m3.x);
      chance = um::saturate(fresnel * glossiness * g);
    }
    if (pure_diffuse) {
      out.dir = hemi;
    } else if (chance > 0 && u2f(r.z) < chance) {
      out.dir = um::reflect(D, rough_normal);
      out.reflectance = um::mk(1.0f);
    } else if (metallic > 0 && u2f(r.w) < metallic) {
      out.dir = um::reflect(D, rough_normal);
    } else {   // a glossy or part-metallic material falling through to its diffuse lobe: second block
      const uint4 r1 = philox4x32_10(pixel, sample, bounce, 1, seed, kPhiloxKey1);
      float s1, c1;
      unit_angle_sincos(u2f(r1.y), &s1, &c1);
      out.dir = cosine_hemisphere(N, u2f(r1.x), s1, c1);
    }
    float ev = ev0;
    if (chance > 0 && chance < 1) ev++;
    if (metallic > 0 && metallic < 1) ev++;
    ev += roughness * (chance + (1 - chance) * metallic);
    ev += (1 - chance) * (1 - metallic);
    out.random_events = ev;
  } else {
    const float ior = m2.y;
    f3 rough_normal = roughness > 0 ? um::normalize(N + roughness * random_direction(ux, sn, cs)) : um::normalize(N);
    float ni_over_nt, cosine;
    f3 outward;
    const float ddn = um::dot(D, rough_normal);
    if (ddn > 0) {
      outward = -rough_normal;
      ni_over_nt = ior;
      cosine = ior * ddn;
    } else {
      outward = rough_normal;
      ni_over_nt = um::div(1.0f, ior);
      cosine = -ddn;
    }
    // Material.Refract (Material.cs:198-210)
    const float dt = um::dot(D, outward);
    const float disc = 1 - ni_over_nt * ni_over_nt * (1 - dt * dt);
    bool refracted = false;
    if (disc > 0) {
      if (u2f(r.z) > schlick(cosine, m3.y, m3.z)) {
        out.dir = ni_over_nt * (D - outward * dt) - outward * um::sqrt(disc);
        refracted = true;
      }
    }
    if (!refracted) {
      out.dir = um::reflect(D, rough_normal);
      out.reflectance = um::mk(1.0f);
    }
    float ev = ev0;
    ev++;
    ev += roughness;
    out.random_events = ev;
  }
  return out;
}

// ---------------------------------------------------------------------------------------
// The reference's OWN random stream (NoiseColor.White): Unity.Mathematics.Random, one xorshift32 state per pixel per
// batch seeded (Seed * 0x8C4CA03F) ^ (index * 0x7383ED49) (SampleBatchJob.cs:91), consumed in call order — a draw
// that the reference skips is not made (RandomSource.cs; draw order: SURVEY.md appendix A.2).  Only the
// thread-per-pixel kernel can follow a sequential stream (RTB_OPT_NOISE = 1): it reproduces the CPU restatement of
// the reference run with its own generator, bit for bit.
// ---------------------------------------------------------------------------------------
struct WhiteNoise {
  uint32_t state;
  __device__ __forceinline__ void init(uint32_t seed) { state = seed; next_state(); }       // Random(uint seed)
  __device__ __forceinline__ uint32_t next_state() {
    const uint32_t t = state;
    state ^= state << 13;
    state ^= state >> 17;
    state ^= state << 5;
    return t;
  }
  __device__ __forceinline__ float next_float() { return u2f(next_state()); }
};

// View.GetRay (View.cs:38-48) after the jitter draw of SampleBatchJob.cs:134: lens disk (2 draws, only with a lens), time (1 draw, always)
__device__ __forceinline__ PathRay camera_ray_white(const rtb_batch_params& p, int cx, int cy, WhiteNoise& rng, float* time) {
  const rtb_view& v = p.view;
  float jx = 0.5f, jy = 0.5f, rdx = 0, rdy = 0;
  if (p.sub_pixel_jitter) { jx = rng.next_float(); jy = rng.next_float(); }
  const float nx = um::div((float)cx + jx, p.size[0]);
  const float ny = um::div((float)cy + jy, p.size[1]);
  if (v.lens_radius != 0) {   // RandomSource.InUnitDisk (RandomSource.cs:40-61)
    const float theta = rng.next_float() * (2 * um::PI - 0) + 0;
    const float radius = um::sqrt(rng.next_float());
    float s, c;
    um::sincos(theta, &s, &c);
    rdx = v.lens_radius * (radius * c);
    rdy = v.lens_radius * (radius * s);
  }
  const f3 offset = v3(v.right) * rdx + v3(v.up) * rdy;
  PathRay ray;
  ray.d = um::normalize(v3(v.lower_left_corner) - offset + nx * v3(v.horizontal) + ny * v3(v.vertical));
  ray.o = v3(v.origin) + offset;
  *time = rng.next_float();   // Ray.Time (View.cs:47): always drawn; read by moving entities only
  return ray;
}

// Material.Scatter (Material.cs:67-173) with the reference's draws in the reference's order
__device__ __forceinline__ ScatterResult scatter_white(float4 m0, float4 m1, float4 m2, float4 m3, f3 D, f3 N, WhiteNoise& rng, float ev0 = 0.0f) {
  ScatterResult out;
  out.reflectance = um::mk(m0.x, m0.y, m0.z);
  const float glossiness = m1.w, metallic = m2.x, roughness = m2.w;
  auto cosine_sample = [&](f3 normal) {          // RandomSource.OnCosineWeightedHemisphere: NextFloat2, x then y
    const float ux = rng.next_float(), uy = rng.next_float();
    float s, c;
    unit_angle_sincos(uy, &s, &c);
    return cosine_hemisphere(normal, ux, s, c);
  };
  if (__float_as_uint(m0.w) == RTB_MATERIAL_STANDARD) {
    const f3 rough_normal = roughness > 0 ? um::normalize(um::lerp(N, cosine_sample(N), roughness)) : N;   // :83
    const float incident_cosine = -um::dot(D, rough_normal);
    const float fresnel = schlick(incident_cosine, m3.y, m3.z);
    const float g = smith_masking_shadowing(D, N, m3.x);
    const float chance = um::saturate(fresnel * glossiness * g);
    if (chance > 0 && rng.next_float() < chance) {                                                          // :91
      out.dir = um::reflect(D, rough_normal);
      out.reflectance = um::mk(1.0f);
    } else if (metallic > 0 && rng.next_float() < metallic) {                                               // :99
      out.dir = um::reflect(D, rough_normal);
    } else {
      out.dir = cosine_sample(N);                                                                           // :107
    }
    float ev = ev0;
    if (chance > 0 && chance < 1) ev++;
    if (metallic > 0 && metallic < 1) ev++;
    ev += roughness * (chance + (1 - chance) * metallic);
    ev += (1 - chance) * (1 - metallic);
    out.random_events = ev;
  } else {
    const float ior = m2.y;
    const float rx = rng.next_float(), ry = rng.next_float();          // NextFloat3Direction: always drawn (:124)
    float s, c;
    unit_angle_sincos(ry, &s, &c);
    const f3 rough_normal = um::normalize(N + roughness * random_direction(rx, s, c));
    float ni_over_nt, cosine;
    f3 outward;
    const float ddn = um::dot(D, rough_normal);
    if (ddn > 0) { outward = -rough_normal; ni_over_nt = ior; cosine = ior * ddn; }
    else { outward = rough_normal; ni_over_nt = um::div(1.0f, ior); cosine = -ddn; }
    const float dt = um::dot(D, outward);
    const float disc = 1 - ni_over_nt * ni_over_nt * (1 - dt * dt);
    bool refracted = false;
    if (disc > 0) {                                                     // Refract succeeded: the Schlick draw is made (:142-143)
      if (rng.next_float() > schlick(cosine, m3.y, m3.z)) {
        out.dir = ni_over_nt * (D - outward * dt) - outward * um::sqrt(disc);
        refracted = true;
      }
    }
    if (!refracted) {
      out.dir = um::reflect(D, rough_normal);
      out.reflectance = um::mk(1.0f);
    }
    out.random_events = (ev0 + 1.0f) + roughness;
  }
  return out;
}

// Texture.SampleColor / SampleScalar for TextureType.Image (Texture.cs:80-89,128-137), texel clamped to the image.
__device__ __forceinline__ const unsigned char* texel_of(const SceneDesc& sd, int image, float tu, float tv) {
  const int4 im = __ldg(sd.tex_images + image);
  int cx = (int)(tu * (float)im.y), cy = (int)(tv * (float)im.z);
  cx = min(max(cx, 0), im.y - 1);
  cy = min(max(cy, 0), im.z - 1);
  return sd.tex_pixels + (size_t)(uint32_t)im.x + ((size_t)cy * (size_t)im.y + (size_t)cx) * (size_t)im.w;
}

__device__ __forceinline__ void derive_material(DevMaterial& x);

// A hit on a material with image textures: Material.Scatter / Emit sample Albedo, Metallic, Glossiness and Emission at
// HitRecord.TexCoords first (Material.cs:70-77,175-179).  The sampled values replace the constants in the material
// registers and the derived Scatter constants are re-evaluated by the function the upload kernel uses.
template <bool SMEM>
__device__ __noinline__ void resolve_textures(const SceneView<SMEM>& sv, const SceneDesc& sd, uint32_t material, float4 prim, f3 o, f3 d,
                                              float4& m0, float4& m1, float4& m2, float4& m3) {
  float tu = 0, tv = 0;                          // "TODO: Texcoord support for primitives" (Entity.cs:107)
  if (prim.w != prim.w && __float_as_uint(prim.y) == 0u && sd.tri_uv) {
    const uint32_t tri = __float_as_uint(prim.x);
    float u = 0, v = 0, t = 0;
    triangle_uvt(sv, tri, o, d, &u, &v, &t);
    const float bx = 1 - u - v;
    const float2 t0 = __ldg(sd.tri_uv + 3 * (size_t)tri), t1 = __ldg(sd.tri_uv + 3 * (size_t)tri + 1), t2 = __ldg(sd.tri_uv + 3 * (size_t)tri + 2);
    tu = um::fma(t2.x, v, um::fma(t1.x, u, t0.x * bx));     // mul(float2x3, barycentricCoords) (HitTests.cs:147)
    tv = um::fma(t2.y, v, um::fma(t1.y, u, t0.y * bx));
  }
  const int4 ti = __ldg(sd.mat_textures + 2 * (size_t)material), tc = __ldg(sd.mat_textures + 2 * (size_t)material + 1);
  DevMaterial x;
  x.albedo[0] = m0.x; x.albedo[1] = m0.y; x.albedo[2] = m0.z; x.type = __float_as_uint(m0.w);
  x.emission[0] = m1.x; x.emission[1] = m1.y; x.emission[2] = m1.z; x.glossiness = m1.w;
  x.metallic = m2.x; x.ior = m2.y; x.perfect_specular = __float_as_uint(m2.z); x.roughness = m2.w;
  x.alpha = m3.x; x.r0 = m3.y; x.one_minus_r0 = m3.z; x.textured = __float_as_uint(m3.w);
  if (ti.x >= 0) {
    const unsigned char* px = texel_of(sd, ti.x, tu, tv);
    for (int k = 0; k < 3; k++) x.albedo[k] = um::div((float)px[k], 255.0f) * x.albedo[k];
  }
  if (ti.y >= 0) {
    const unsigned char* px = texel_of(sd, ti.y, tu, tv);
    for (int k = 0; k < 3; k++) x.emission[k] = um::div((float)px[k], 255.0f) * x.emission[k];
  }
  if (ti.z >= 0) x.glossiness = um::div((float)texel_of(sd, ti.z, tu, tv)[tc.x], 255.0f) * x.glossiness;
  if (ti.w >= 0) x.metallic = um::div((float)texel_of(sd, ti.w, tu, tv)[tc.y], 255.0f) * x.metallic;
  if (x.type == RTB_MATERIAL_DIELECTRIC) x.ior = m2.y;      // IndexOfRefraction is not a texture
  derive_material(x);
  m0 = make_float4(x.albedo[0], x.albedo[1], x.albedo[2], m0.w);
  m1 = make_float4(x.emission[0], x.emission[1], x.emission[2], x.glossiness);
  m2 = make_float4(x.metallic, x.ior, m2.z, x.roughness);
  m3 = make_float4(x.alpha, x.r0, x.one_minus_r0, m3.w);
}

// DevMaterial's derived constants from (type, glossiness, metallic, ior): the per-material part of Material.Scatter.
// Fills DevMaterial's derived constants on the device (one thread per material, at upload).
__global__ void derive_materials_kernel(DevMaterial* m, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DevMaterial x = m[i];
  derive_material(x);
  m[i] = x;
}
__device__ __forceinline__ void derive_material(DevMaterial& x) {
  if (x.type == RTB_MATERIAL_STANDARD) {
    x.roughness = um::pow2(1 - x.glossiness);
    x.ior = um::lerp(1.5f, 1.1f, x.metallic);        // PlasticIor, MetalIor (Material.cs:18-19)
    x.alpha = roughness_to_alpha(x.roughness);
  } else {
    x.roughness = 1 - x.glossiness;
    x.alpha = 0;
  }
  x.r0 = schlick_r0(x.ior);
  x.one_minus_r0 = 1 - x.r0;
}

// 2^-depth exactly as repeated halving produces it (normal, denormal, then 0)
__device__ __forceinline__ float pow2_neg(uint32_t depth) {
  if (depth <= 126u) return __uint_as_float((127u - depth) << 23);
  if (depth <= 149u) return __uint_as_float(1u << (149u - depth));
  return 0.0f;
}

// Cubemap.Sample (Texture.cs:171-210): major axis (lowest index on ties), face uv, nearest texel, halves -> floats
__device__ __forceinline__ f3 cubemap_sample(const uint16_t* faces, int fw, int fh, f3 v) {
  if (!faces) return um::mk(0.0f);
  const float ax = um::abs(v.x), ay = um::abs(v.y), az = um::abs(v.z);
  const float max_distance = um::max(um::max(um::max(ax, ay), az), 0.0f);
  const int lane = max_distance == ax ? 0 : (max_distance == ay ? 1 : (max_distance == az ? 2 : 3));
  float major = ax, comp = v.x, u, w;
  if (lane == 0) { const bool positive = v.x >= 0; u = positive ? -v.z : v.z; w = -v.y; }
  else if (lane == 1) { const bool positive = v.y >= 0; major = ay; comp = v.y; u = v.x; w = positive ? v.z : -v.z; }
  else { const bool positive = v.z >= 0; major = az; comp = v.z; u = positive ? v.x : -v.x; w = -v.y; }
  const bool positive = comp >= 0;
  u = um::div(u, major);
  w = um::div(w, major);
  const int cx = min((int)((u + 1) * (float)(fw / 2)), fw - 1);
  const int cy = min((int)((w + 1) * (float)(fh / 2)), fh - 1);
  const uint16_t* px = faces + ((size_t)((lane > 2 ? 2 : lane) * 2 + (positive ? 0 : 1)) * (size_t)fw * (size_t)fh + (size_t)cy * (size_t)fw + (size_t)cx) * 4;
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(px));
  return um::mk(um::half_to_float((uint16_t)(raw.x & 0xffffu)), um::half_to_float((uint16_t)(raw.x >> 16)),
                um::half_to_float((uint16_t)(raw.y & 0xffffu)));
}

// Sky (SampleBatchJob.cs:348-374; Environment.cs)
__device__ __forceinline__ f3 sky_color(const rtb_environment& e, const SceneDesc& sd, f3 d) {
  if (e.sky_type == RTB_SKY_GRADIENT)
    return um::lerp(v3(e.sky_bottom_color), v3(e.sky_top_color), 0.5f * (d.y + 1));
  if (e.sky_type == RTB_SKY_CUBEMAP) return cubemap_sample(sd.sky_faces, sd.sky_w, sd.sky_h, d);
  return um::mk(0.0f);
}

// Per-pixel prologue (SampleBatchJob.cs:72-126): how many samples this batch takes.
__device__ __forceinline__ uint32_t samples_to_accumulate(const rtb_batch_params& p, float in_w, float in_weight,
                                                          float* sample_count_weight) {
  int sample_count = (int)in_w;
  float w = um::div(in_weight, (float)sample_count);
  *sample_count_weight = w;
  if (w == 0) return p.sample_count_range[0];
  float normalized = um::saturate(um::unlerp(p.sample_count_weight_extrema[0], p.sample_count_weight_extrema[1], w));
  return (uint32_t)um::round(um::lerp((float)p.sample_count_range[0], (float)p.sample_count_range[1], normalized));
}

}  // namespace rtbk
