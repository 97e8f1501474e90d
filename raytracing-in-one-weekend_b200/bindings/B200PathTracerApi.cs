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
		// rtb_option (include/rtb.h): Counters = 1, Kernel = 2, CancelChunkRows = 3, LeafSpheres = 4, AlwaysWalkChains = 5, HostAccess = 6, Noise = 7
		[DllImport(Lib)] public static extern RtbStatus rtb_set_option(IntPtr ctx, int option, long value);
		[DllImport(Lib)] public static extern RtbStatus rtb_last_kernel_ms(IntPtr ctx, out float ms);
		[DllImport(Lib)] public static extern RtbStatus rtb_last_batch_in_place(IntPtr ctx, out int inPlace);
		public static string LastError(IntPtr ctx) => Marshal.PtrToStringAnsi(rtb_last_error(ctx));
	}

	// Replaces `sampleBatchJob.Schedule(totalBufferSize, 1, dep)` (Raytracer.cs:730-736).  Non-Burst IJob, like
	// OpenImageDenoiseJob (DenoiseJobs.cs:9-39): the one Execute() blocks until the out* arrays are written.
	public unsafe struct B200SampleBatchJob : IJob
	{
		[NativeDisableUnsafePtrRestriction] public IntPtr Context;
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
			var status = Api.rtb_sample_batch(Context, &p, &b, (bool*) CancellationToken.GetUnsafePtrWithoutChecks());
			if (status != RtbStatus.Ok && status != RtbStatus.Cancelled)
				UnityEngine.Debug.LogError($"rtb_sample_batch: {status} {Api.LastError(Context)}");   // same convention as Raytracer.cs:341-365
		}
	}

	// Flattens the host's pointer graph (BvhNode*/Entity*/Material*, all in contiguous NativeList/NativeArray:
	// Raytracer.cs:155,160-162) into the index-based arrays rtb_upload_scene takes.  Called from RebuildWorld
	// (Raytracer.cs:1167-1183) after BuildRuntimeBvhJob.
	public static unsafe class SceneFlattener
	{
		public static RtbStatus Upload(IntPtr ctx, NativeArray<Runtime.BvhNode> nodes, NativeList<Runtime.Entity> entities, NativeList<Runtime.Material> materials)
		{
			var nodeBase = (Runtime.BvhNode*) nodes.GetUnsafeReadOnlyPtr();
			var entityBase = (Runtime.Entity*) entities.GetUnsafeReadOnlyPtr();
			var materialBase = (Runtime.Material*) materials.GetUnsafeReadOnlyPtr();
			var outNodes = new NativeArray<RtbBvhNode>(nodes.Length, Allocator.Temp);
			var outSpheres = new NativeArray<RtbSphere>(entities.Length, Allocator.Temp);
			var outMaterials = new NativeArray<RtbMaterial>(materials.Length, Allocator.Temp);
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
				Runtime.Entity e = entities[i];        // EntityType.Sphere only (anything else: keep the Burst path)
				var s = (Runtime.EntityTypes.Sphere*) e.Content;
				outSpheres[i] = new RtbSphere { Center = e.OriginTransform.pos, Radius = s->Radius, Material = (uint) (e.Material - materialBase) };
			}
			for (int i = 0; i < materials.Length; i++)
			{
				Runtime.Material m = materials[i];     // constant textures only (Texture.cs:50-59)
				outMaterials[i] = new RtbMaterial
				{
					Type = (uint) m.Type, Albedo = m.Albedo.MainColor, Emission = m.Emission.MainColor,
					Glossiness = m.Glossiness.MainColor.x, Metallic = m.Metallic.MainColor.x, IndexOfRefraction = m.IndexOfRefraction,
				};
			}
			return Api.rtb_upload_scene(ctx, (RtbSphere*) outSpheres.GetUnsafeReadOnlyPtr(), (UIntPtr) outSpheres.Length,
				(RtbMaterial*) outMaterials.GetUnsafeReadOnlyPtr(), (UIntPtr) outMaterials.Length,
				(RtbBvhNode*) outNodes.GetUnsafeReadOnlyPtr(), (UIntPtr) outNodes.Length);
		}
	}
}
#endif
