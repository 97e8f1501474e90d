// B200PathTracerApi.cs — the P/Invoke binding a maintainer of renaudbedard/raytracing-in-one-weekend
// would drop into Assets/ThirdParty/B200PathTracer/ to put librtb.so (include/rtb.h) behind
// Raytracer.ScheduleSample.  NOT compiled in this repository (no Unity / .NET toolchain in the image);
// it mirrors include/rtb.h field for field and follows the style of the repo's own native binding,
// Assets/ThirdParty/nVidia OptiX Denoiser/OptixApi.cs:154-251 ([DllImport] externs on IntPtr handles)
// and the job shape of Runtime/Jobs/DenoiseJobs.cs:10-39 (a non-Burst IJob making a blocking native call).
//
// Python twin used by this repo's tests: raytracing-in-one-weekend_b200/plugin.py.
#if ENABLE_B200_PATHTRACER
using System;
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Jobs;
using Unity.Mathematics;

namespace B200PathTracer
{
	public enum RtbStatus { Ok = 0, InvalidArgument = 1, NoScene = 2, Cancelled = 3, Unsupported = 4, OutOfMemory = 5, Cuda = 100 }

	[StructLayout(LayoutKind.Sequential)] public struct RtbSphere { public float3 Center; public float Radius; public uint Material; uint r0, r1, r2; }                       // 32 B
	[StructLayout(LayoutKind.Sequential)] public struct RtbMaterial { public uint Type; public float3 Albedo, Emission; public float Glossiness, Metallic, IndexOfRefraction; uint r0, r1; } // 48 B
	[StructLayout(LayoutKind.Sequential)] public struct RtbBvhNode { public float3 Min, Max; public int Left, Right, FirstEntity, EntityCount; }                           // 40 B
	[StructLayout(LayoutKind.Sequential)] public struct RtbEnvironment { public uint SkyType; public float3 SkyBottomColor, SkyTopColor; }

	[StructLayout(LayoutKind.Sequential)]
	public struct RtbBatchParams            // == the uniform fields of SampleBatchJob (SampleBatchJob.cs:25-39)
	{
		public float2 Size; public int SliceOffset, SliceDivider; public uint Seed;
		public Runtime.View View;           // 88 bytes, blittable as is (View.cs:8-14)
		public RtbEnvironment Environment;
		public uint2 SampleCountRange; public int TraceDepth; public uint SubPixelJitter;
		public float2 SampleCountWeightExtrema;
		public int RowBegin, RowEnd;        // extension: row-tile sharding (0,0 = all rows)
	}

	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtbBatchBuffers    // == the NativeArray fields of SampleBatchJob (SampleBatchJob.cs:41-51)
	{
		public float4* InColor; public float* InSampleCountWeight; public float3* InNormal; public float3* InAlbedo;
		public float4* OutColor; public float* OutSampleCountWeight; public float3* OutNormal; public float3* OutAlbedo;
		public Diagnostics* OutDiagnostics; // FULL_DIAGNOSTICS layout: 4 floats (Raytracer.cs:54-64)
	}

	// rtb_triangle: EntityTypes/Triangle.cs:10-11 (Data columns, Normals columns), world space
	[StructLayout(LayoutKind.Sequential)] public struct RtbTriangle
	{
		public float3 Edge2, Edge1, V0;     // Triangle.Data[0], Data[1], Data[2]
		public float3 N0, N1, N2;           // Triangle.Normals columns
		public uint Material, Reserved;
	}
	// rtb_entity: one element of bvhEntities (BvhNodeData.cs:157-160): Entity.Type + index of Entity.Content
	[StructLayout(LayoutKind.Sequential)] public struct RtbEntity { public uint Type; /* EntityType: 1 Sphere, 2 Rect, 3 Box, 4 Triangle; | 0x100: Index is into the placed-entity array */ public uint Index; }
	// rtb_placed_entity: Entity.cs:24-37 with the content inlined (rotated / moving entities, Rect, Box)
	[StructLayout(LayoutKind.Sequential)] public struct RtbPlacedEntity
	{
		public uint Type, Material, Moving, Reserved;       // Entity.Type, Material - materialBuffer.ptr, Entity.Moving
		public quaternion Rotation; public float3 Position;  // Entity.OriginTransform
		public float3 DestinationOffset; public float2 TimeRange;
		public float3 Size; public float Reserved2;          // Sphere: (Radius, -, -); Rect: ctor size.xy; Box: ctor size.xyz
	}

	// rtb_image / rtb_material_textures: TextureType.Image textures (Texture.cs:22-48) of the mesh materials (Raytracer.cs:1210-1267)
	[StructLayout(LayoutKind.Sequential)] public unsafe struct RtbImage { public byte* Pixels; public int Width, Height, PixelStride, Reserved; }
	[StructLayout(LayoutKind.Sequential)] public struct RtbMaterialTextures
	{
		public int AlbedoImage, EmissionImage, GlossinessImage, MetallicImage;   // index into the image array or -1 (TextureType.Constant)
		public int GlossinessChannel, MetallicChannel, Reserved0, Reserved1;     // Texture.ScalarValueChannel
	}

	public static unsafe class Api
	{
		const string Lib = "rtb";           // librtb.so / rtb.dll next to the other native plugins
		[DllImport(Lib)] public static extern int rtb_abi_version();
		[DllImport(Lib)] public static extern RtbStatus rtb_create(int device, out IntPtr ctx);
		[DllImport(Lib)] public static extern RtbStatus rtb_destroy(IntPtr ctx);
		[DllImport(Lib)] public static extern IntPtr rtb_last_error(IntPtr ctx);
		[DllImport(Lib)] public static extern RtbStatus rtb_upload_scene(IntPtr ctx, RtbSphere* spheres, UIntPtr sphereCount,
			RtbMaterial* materials, UIntPtr materialCount, RtbBvhNode* nodes, UIntPtr nodeCount);
		// the topology the device walks under RTB_OPT_RETREE (host-side, no device): same leaves as `nodes`, another tree above them
		[DllImport(Lib)] public static extern RtbStatus rtb_retree_bvh(RtbBvhNode* nodes, UIntPtr nodeCount, RtbBvhNode* outNodes, UIntPtr capacity, out UIntPtr outCount);
		[DllImport(Lib)] public static extern RtbStatus rtb_upload_world(IntPtr ctx, RtbEntity* entities, UIntPtr entityCount,
			RtbSphere* spheres, UIntPtr sphereCount, RtbTriangle* triangles, UIntPtr triangleCount,
			RtbMaterial* materials, UIntPtr materialCount, RtbBvhNode* nodes, UIntPtr nodeCount);
		[DllImport(Lib)] public static extern RtbStatus rtb_upload_placed_world(IntPtr ctx, RtbEntity* entities, UIntPtr entityCount,
			RtbSphere* spheres, UIntPtr sphereCount, RtbTriangle* triangles, UIntPtr triangleCount, RtbPlacedEntity* placed, UIntPtr placedCount,
			RtbMaterial* materials, UIntPtr materialCount, RtbBvhNode* nodes, UIntPtr nodeCount);
		[DllImport(Lib)] public static extern RtbStatus rtb_upload_textures(IntPtr ctx, RtbImage* images, UIntPtr imageCount,
			RtbMaterialTextures* materialTextures, UIntPtr materialCount, float* triangleUvs, UIntPtr triangleCount);
		// Environment.SkyCubemap: cubemap.GetPixelData<byte>(0, CubemapFace.PositiveX) of an R16G16B16A16_SFloat cubemap (Texture.cs:155-167)
		[DllImport(Lib)] public static extern RtbStatus rtb_upload_sky_cubemap(IntPtr ctx, ushort* halfRgba, int faceWidth, int faceHeight);
		[DllImport(Lib)] public static extern RtbStatus rtb_sample_batch(IntPtr ctx, RtbBatchParams* p, RtbBatchBuffers* hostBuffers, bool* cancel);
		[DllImport(Lib)] public static extern RtbStatus rtb_register_host_buffer(IntPtr ctx, void* ptr, UIntPtr bytes);
		[DllImport(Lib)] public static extern RtbStatus rtb_unregister_host_buffer(IntPtr ctx, void* ptr);
		[DllImport(Lib)] public static extern RtbStatus rtb_combine_device(IntPtr ctx, int width, int height, int debugMode, int ldrAlbedo,
			float* color4, float* normal3, float* albedo3, float* outColor3, float* outNormal3, float* outAlbedo3, IntPtr cudaStream);
		[DllImport(Lib)] public static extern RtbStatus rtb_finalize_device(IntPtr ctx, int width, int height,
			float* color3, float* normal3, float* albedo3, uint* outColorRgba, uint* outNormalRgba, uint* outAlbedoRgba, IntPtr cudaStream);
		// ---- one host, N GPUs: the same call site, every GPU of the box (include/rtb.h "one host, N GPUs") ----
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_create(int* devices, int deviceCount, out IntPtr multi);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_destroy(IntPtr multi);
		[DllImport(Lib)] public static extern int rtb_multi_device_count(IntPtr multi);
		[DllImport(Lib)] public static extern IntPtr rtb_multi_context(IntPtr multi, int index);
		[DllImport(Lib)] public static extern IntPtr rtb_multi_last_error(IntPtr multi);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_set_option(IntPtr multi, int option, long value);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_upload_scene(IntPtr multi, RtbSphere* spheres, UIntPtr sphereCount,
			RtbMaterial* materials, UIntPtr materialCount, RtbBvhNode* nodes, UIntPtr nodeCount);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_upload_placed_world(IntPtr multi, RtbEntity* entities, UIntPtr entityCount,
			RtbSphere* spheres, UIntPtr sphereCount, RtbTriangle* triangles, UIntPtr triangleCount, RtbPlacedEntity* placed, UIntPtr placedCount,
			RtbMaterial* materials, UIntPtr materialCount, RtbBvhNode* nodes, UIntPtr nodeCount);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_upload_textures(IntPtr multi, RtbImage* images, UIntPtr imageCount,
			RtbMaterialTextures* materialTextures, UIntPtr materialCount, float* triangleUvs, UIntPtr triangleCount);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_upload_sky_cubemap(IntPtr multi, ushort* halfRgba, int faceWidth, int faceHeight);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_register_host_buffer(IntPtr multi, void* ptr, UIntPtr bytes);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_unregister_host_buffer(IntPtr multi, void* ptr);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_sample_batch(IntPtr multi, RtbBatchParams* p, RtbBatchBuffers* hostBuffers, bool* cancel);
		[DllImport(Lib)] public static extern RtbStatus rtb_multi_get_tiles(IntPtr multi, int* outBounds, float* outKernelMs);
		// rtb_option (include/rtb.h): Counters = 1, Kernel = 2, LeafSpheres = 4, AlwaysWalkChains = 5, HostAccess = 6, Noise = 7, BalanceTiles = 8, Math = 9, Retree = 10
		[DllImport(Lib)] public static extern RtbStatus rtb_set_option(IntPtr ctx, int option, long value);
		[DllImport(Lib)] public static extern RtbStatus rtb_last_kernel_ms(IntPtr ctx, out float ms);
		[DllImport(Lib)] public static extern RtbStatus rtb_last_batch_in_place(IntPtr ctx, out int inPlace);
		public static string LastError(IntPtr ctx) => Marshal.PtrToStringAnsi(rtb_last_error(ctx));
	}

	// Replaces `sampleBatchJob.Schedule(totalBufferSize, 1, dep)` (Raytracer.cs:730-736).  Non-Burst IJob, like
	// OpenImageDenoiseJob (DenoiseJobs.cs:9-39): the one Execute() blocks until the out* arrays are written.
	public unsafe struct B200SampleBatchJob : IJob
	{
		[NativeDisableUnsafePtrRestriction] public IntPtr Context;   // rtb_ctx*, or rtb_multi* with Multi = true (every GPU of the box)
		public bool Multi;
		[ReadOnly] public NativeReference<bool> CancellationToken;
		public RtbBatchParams Params;
		[ReadOnly] public NativeArray<float4> InputColor;
		[ReadOnly] public NativeArray<float> InputSampleCountWeight;
		[ReadOnly] public NativeArray<float3> InputNormal, InputAlbedo;
		[WriteOnly] public NativeArray<float4> OutputColor;
		[WriteOnly] public NativeArray<float> OutputSampleCountWeight;
		[WriteOnly] public NativeArray<float3> OutputNormal, OutputAlbedo;
		[WriteOnly] public NativeArray<Diagnostics> OutputDiagnostics;

		public void Execute()
		{
			var p = Params;
			var b = new RtbBatchBuffers
			{
				InColor = (float4*) InputColor.GetUnsafeReadOnlyPtr(), InSampleCountWeight = (float*) InputSampleCountWeight.GetUnsafeReadOnlyPtr(),
				InNormal = (float3*) InputNormal.GetUnsafeReadOnlyPtr(), InAlbedo = (float3*) InputAlbedo.GetUnsafeReadOnlyPtr(),
				OutColor = (float4*) OutputColor.GetUnsafePtr(), OutSampleCountWeight = (float*) OutputSampleCountWeight.GetUnsafePtr(),
				OutNormal = (float3*) OutputNormal.GetUnsafePtr(), OutAlbedo = (float3*) OutputAlbedo.GetUnsafePtr(),
				OutDiagnostics = (Diagnostics*) OutputDiagnostics.GetUnsafePtr(),
			};
			// the token is polled INSIDE the one kernel launch (rtb.h): passing it costs nothing, setting it ends the call within a millisecond
			var token = (bool*) CancellationToken.GetUnsafePtrWithoutChecks();
			var status = Multi ? Api.rtb_multi_sample_batch(Context, &p, &b, token) : Api.rtb_sample_batch(Context, &p, &b, token);
			if (status != RtbStatus.Ok && status != RtbStatus.Cancelled)      // same convention as Raytracer.cs:341-365
				UnityEngine.Debug.LogError($"rtb_sample_batch: {status} {Marshal.PtrToStringAnsi(Multi ? Api.rtb_multi_last_error(Context) : Api.rtb_last_error(Context))}");
		}
	}

	// Flattens the host's pointer graph (BvhNode*/Entity*/Material*, all in contiguous NativeList/NativeArray:
	// Raytracer.cs:155,160-162) into the index-based arrays the rtb_upload_* calls take.  Called from RebuildWorld
	// (Raytracer.cs:1167-1183) after BuildRuntimeBvhJob.  Covers everything the plugin renders: what the host produces at
	// HEAD — mesh triangles with image textures (AddMeshRuntimeEntitiesJob.cs; Raytracer.cs:1185-1304) — and the primitive
	// entity kinds of the older scenes (Sphere, Rect, Box; rotated or moving).  Anything else returns Unsupported so that
	// the caller keeps the Burst job for that world.
	public static unsafe class SceneFlattener
	{
		static bool IsIdentity(quaternion q) => q.value.x == 0 && q.value.y == 0 && q.value.z == 0 && q.value.w == 1;

		// ctx: rtb_ctx* (one GPU) or, with multi = true, rtb_multi* (every GPU of the box gets the world)
		public static RtbStatus Upload(IntPtr ctx, bool multi, NativeArray<Runtime.BvhNode> nodes, NativeList<Runtime.Entity> entities,
			NativeList<Runtime.Material> materials)
		{
			var nodeBase = (Runtime.BvhNode*) nodes.GetUnsafeReadOnlyPtr();
			var entityBase = (Runtime.Entity*) entities.GetUnsafeReadOnlyPtr();
			var materialBase = (Runtime.Material*) materials.GetUnsafeReadOnlyPtr();
			var outNodes = new NativeArray<RtbBvhNode>(nodes.Length, Allocator.Temp);
			var outEntities = new NativeArray<RtbEntity>(entities.Length, Allocator.Temp);
			var spheres = new NativeList<RtbSphere>(entities.Length, Allocator.Temp);
			var triangles = new NativeList<RtbTriangle>(entities.Length, Allocator.Temp);
			var triangleUvs = new NativeList<float2>(entities.Length * 3, Allocator.Temp);
			var placed = new NativeList<RtbPlacedEntity>(16, Allocator.Temp);
			var outMaterials = new NativeArray<RtbMaterial>(materials.Length, Allocator.Temp);
			var materialTextures = new NativeArray<RtbMaterialTextures>(materials.Length, Allocator.Temp);
			var images = new NativeList<RtbImage>(16, Allocator.Temp);
			bool anyImage = false;

			for (int i = 0; i < nodes.Length; i++)
			{
				Runtime.BvhNode n = nodes[i];
				outNodes[i] = new RtbBvhNode
				{
					Min = n.Bounds.Min, Max = n.Bounds.Max,
					Left = n.Left != null ? (int) (n.Left - nodeBase) : -1, Right = n.Right != null ? (int) (n.Right - nodeBase) : -1,
					FirstEntity = n.IsLeaf ? (int) (n.EntitiesStart - entityBase) : -1, EntityCount = n.EntityCount,
				};
			}

			for (int i = 0; i < entities.Length; i++)
			{
				Runtime.Entity e = entities[i];
				uint material = (uint) (e.Material - materialBase);
				switch (e.Type)
				{
					case Runtime.EntityType.Triangle:      // always world space (Entity.cs:92-93): no transform crosses
					{
						var t = (Runtime.EntityTypes.Triangle*) e.Content;
						outEntities[i] = new RtbEntity { Type = 4, Index = (uint) triangles.Length };
						triangles.Add(new RtbTriangle
						{
							Edge2 = t->Data.c0, Edge1 = t->Data.c1, V0 = t->Data.c2,
							N0 = t->Normals.c0, N1 = t->Normals.c1, N2 = t->Normals.c2, Material = material,
						});
						triangleUvs.Add(t->TextureCoordinates.c0); triangleUvs.Add(t->TextureCoordinates.c1); triangleUvs.Add(t->TextureCoordinates.c2);
						break;
					}
					case Runtime.EntityType.Sphere when !e.Moving && IsIdentity(e.OriginTransform.rot):
					{
						var sp = (Runtime.EntityTypes.Sphere*) e.Content;   // the plain sphere: translation only
						outEntities[i] = new RtbEntity { Type = 1, Index = (uint) spheres.Length };
						spheres.Add(new RtbSphere { Center = e.OriginTransform.pos, Radius = sp->Radius, Material = material });
						break;
					}
					case Runtime.EntityType.Sphere:
					case Runtime.EntityType.Rect:
					case Runtime.EntityType.Box:
					{
						// the full Entity record (rotation, motion, content): rtb_placed_entity.  `size` is what the content's
						// constructor took: Sphere radius; Rect size = To - From (Rect.cs:11-15); Box size = 2 * Extents (Box.cs:11-15)
						float3 size = default;
						if (e.Type == Runtime.EntityType.Sphere) size.x = ((Runtime.EntityTypes.Sphere*) e.Content)->Radius;
						else if (e.Type == Runtime.EntityType.Rect) { var r = (Runtime.EntityTypes.Rect*) e.Content; size.xy = r->To - r->From; }
						else size = ((Runtime.EntityTypes.Box*) e.Content)->Extents * 2;
						outEntities[i] = new RtbEntity { Type = (uint) e.Type | 0x100u, Index = (uint) placed.Length };
						placed.Add(new RtbPlacedEntity
						{
							Type = (uint) e.Type, Material = material, Moving = e.Moving ? 1u : 0u,
							Rotation = e.OriginTransform.rot, Position = e.OriginTransform.pos,
							DestinationOffset = e.DestinationOffset, TimeRange = e.TimeRange, Size = size,
						});
						break;
					}
					default:
						return RtbStatus.Unsupported;
				}
			}

			// materials: constants go into rtb_material, TextureType.Image textures (Texture.cs:80-89,128-137) into the image list
			int ImageOf(Runtime.Texture t)
			{
				if (t.Type != Runtime.TextureType.Image) return -1;
				anyImage = true;
				for (int k = 0; k < images.Length; k++)
					if (images[k].Pixels == t.ImagePointer) return k;
				images.Add(new RtbImage { Pixels = t.ImagePointer, Width = t.ImageSize.x, Height = t.ImageSize.y, PixelStride = t.PixelStride });
				return images.Length - 1;
			}
			static bool Supported(Runtime.Texture t) => t.Type == Runtime.TextureType.Constant || t.Type == Runtime.TextureType.ConstantScalar ||
				t.Type == Runtime.TextureType.Image || t.Type == Runtime.TextureType.None;
			// a scalar texture's constant: ConstantScalar keeps it in Parameter (Texture.cs:56-58), the others in MainColor[channel]
			static float Scalar(Runtime.Texture t) => t.Type == Runtime.TextureType.ConstantScalar ? t.Parameter : t.MainColor[t.ScalarValueChannel];
			for (int i = 0; i < materials.Length; i++)
			{
				Runtime.Material m = materials[i];
				if (!Supported(m.Albedo) || !Supported(m.Emission) || !Supported(m.Glossiness) || !Supported(m.Metallic))
					return RtbStatus.Unsupported;          // checker / Perlin textures: commented out in the reference, not carried
				outMaterials[i] = new RtbMaterial
				{
					Type = (uint) m.Type, Albedo = m.Albedo.MainColor, Emission = m.Emission.MainColor,
					Glossiness = Scalar(m.Glossiness), Metallic = Scalar(m.Metallic),
					IndexOfRefraction = m.Type == Runtime.MaterialType.ProbabilisticVolume ? m.Density : m.IndexOfRefraction,   // Material.parameter
				};
				materialTextures[i] = new RtbMaterialTextures
				{
					AlbedoImage = ImageOf(m.Albedo), EmissionImage = ImageOf(m.Emission),
					GlossinessImage = ImageOf(m.Glossiness), MetallicImage = ImageOf(m.Metallic),
					GlossinessChannel = m.Glossiness.ScalarValueChannel, MetallicChannel = m.Metallic.ScalarValueChannel,
				};
			}

			RtbStatus status;
			var pe = (RtbEntity*) outEntities.GetUnsafeReadOnlyPtr(); var ps = (RtbSphere*) spheres.GetUnsafeReadOnlyPtr();
			var pt = (RtbTriangle*) triangles.GetUnsafeReadOnlyPtr(); var pp = (RtbPlacedEntity*) placed.GetUnsafeReadOnlyPtr();
			var pm = (RtbMaterial*) outMaterials.GetUnsafeReadOnlyPtr(); var pn = (RtbBvhNode*) outNodes.GetUnsafeReadOnlyPtr();
			status = multi
				? Api.rtb_multi_upload_placed_world(ctx, pe, (UIntPtr) outEntities.Length, ps, (UIntPtr) spheres.Length, pt, (UIntPtr) triangles.Length,
					pp, (UIntPtr) placed.Length, pm, (UIntPtr) outMaterials.Length, pn, (UIntPtr) outNodes.Length)
				: Api.rtb_upload_placed_world(ctx, pe, (UIntPtr) outEntities.Length, ps, (UIntPtr) spheres.Length, pt, (UIntPtr) triangles.Length,
					pp, (UIntPtr) placed.Length, pm, (UIntPtr) outMaterials.Length, pn, (UIntPtr) outNodes.Length);
			if (status != RtbStatus.Ok || !anyImage) return status;
			var pi = (RtbImage*) images.GetUnsafeReadOnlyPtr(); var pmt = (RtbMaterialTextures*) materialTextures.GetUnsafeReadOnlyPtr();
			var puv = (float*) triangleUvs.GetUnsafeReadOnlyPtr();
			return multi
				? Api.rtb_multi_upload_textures(ctx, pi, (UIntPtr) images.Length, pmt, (UIntPtr) materialTextures.Length, puv, (UIntPtr) triangles.Length)
				: Api.rtb_upload_textures(ctx, pi, (UIntPtr) images.Length, pmt, (UIntPtr) materialTextures.Length, puv, (UIntPtr) triangles.Length);
		}

		// Environment.SkyCubemap (Raytracer.cs:663-665): the six faces of the R16G16B16A16_SFloat cubemap, +X -X +Y -Y +Z -Z
		public static RtbStatus UploadSky(IntPtr ctx, bool multi, UnityEngine.Cubemap cubemap)
		{
			int w = cubemap.width, h = cubemap.height;
			var faces = new NativeArray<ushort>(6 * w * h * 4, Allocator.Temp);
			for (int f = 0; f < 6; f++)
				NativeArray<ushort>.Copy(cubemap.GetPixelData<ushort>(0, (UnityEngine.CubemapFace) f), 0, faces, f * w * h * 4, w * h * 4);
			var p = (ushort*) faces.GetUnsafeReadOnlyPtr();
			return multi ? Api.rtb_multi_upload_sky_cubemap(ctx, p, w, h) : Api.rtb_upload_sky_cubemap(ctx, p, w, h);
		}
	}
}
#endif
