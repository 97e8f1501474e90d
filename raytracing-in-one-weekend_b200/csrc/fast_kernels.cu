// fast_kernels.cu — the opt-in fast-arithmetic build of the lean sphere megakernel (RTB_OPT_MATH = 1).
//
// The reference compiles its job with [BurstCompile(FloatPrecision.Medium, FloatMode.Fast)] (SampleBatchJob.cs:16):
// contraction, reassociation and 3.5-ulp transcendentals are allowed there.  The product build (plugin.cu) fixes one
// strict evaluation so that every path decision can be checked bit for bit against the CPU oracle; THIS translation
// unit is the same kernel source compiled the way the reference's own compiler is allowed to: -fmad=true, MUFU
// approximations for sqrt / divide / sincos / log (include/rtb/umath.h, RTB_FAST_MATH) and the slab test as one FMA per
// plane.  Its images agree with the parity build statistically (same estimator, same Philox draws, a few decisions in
// a million differ); tools/fast_math_report.py reports mean / RMSE / p99.9 of the per-pixel difference.  The headline
// numbers are the parity build's; bench.py reports this build beside them as `value_fast`.
//
// Same source, separate namespace: the kernels are templates, and two instantiations with the same mangled name in
// one library would be merged by the linker.
#define RTB_FAST_MATH 1
#define RTB_ROOT_SKIP 1     // no slab test of the root's own box on a re-built tree (it decides nothing: retree.hpp); measured on this build 98.8 -> 97.2 ms (the parity build: 114.0 -> 114.3, kept)
#define rtbk rtbk_fast
#include "sample_kernels.cuh"

namespace rtbk_fast {

template <bool SMEM, int FLAVOR>
cudaError_t launch(const BatchArgs& a, unsigned grid, size_t smem, int max_smem_optin, cudaStream_t stream) {
  auto kernel = sample_megakernel<SMEM, false, FLAVOR>;
  static bool attr_set = false;          // per process and instantiation; cudaFuncSetAttribute applies to every device's copy
  cudaError_t e = cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);
  static int attr_dev_mask = 0;
  if (!attr_set || !(attr_dev_mask & (1 << dev))) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    attr_set = true;
    attr_dev_mask |= 1 << dev;
  }
  kernel<<<grid, mega_block(FLAVOR), smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace rtbk_fast

// `args`: the parity build's rtbk::BatchArgs (same layout: same header, same compiler) passed as bytes.
// `flavor`: kFlavorSpheres (trees of small leaves) or kFlavorChains (collapsed or big leaves: linear hit lists).
extern "C" __attribute__((visibility("hidden"))) int rtb_fast_launch_spheres(const void* args, size_t args_bytes, int scene_in_smem,
                                                                             int flavor, unsigned grid, size_t smem,
                                                                             int max_smem_optin, void* stream) {
  using namespace rtbk_fast;
  BatchArgs a;
  if (args_bytes != sizeof a || (flavor != kFlavorSpheres && flavor != kFlavorChains)) return (int)cudaErrorInvalidValue;
  memcpy(&a, args, sizeof a);
  cudaStream_t s = (cudaStream_t)stream;
  if (flavor == kFlavorSpheres)
    return (int)(scene_in_smem ? launch<true, kFlavorSpheres>(a, grid, smem, max_smem_optin, s) : launch<false, kFlavorSpheres>(a, grid, smem, max_smem_optin, s));
  return (int)(scene_in_smem ? launch<true, kFlavorChains>(a, grid, smem, max_smem_optin, s) : launch<false, kFlavorChains>(a, grid, smem, max_smem_optin, s));
}
