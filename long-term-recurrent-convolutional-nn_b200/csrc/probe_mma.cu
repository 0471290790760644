// probe_mma.cu -- diagnostics: issue-to-completion cost of a chain of tcgen05.mma (kind::f16, bf16, SS mode, K-major
// SWIZZLE_128B operands resident in smem) accumulating into one TMEM tile, for a given M x N.  One CTA per SM.
#include "kernels.cuh"
#include "sm100_ptx.cuh"

namespace lrcn {
using namespace ptx;

// commit_every > 0: a tcgen05.commit (to a scratch mbarrier with a huge count) after every commit_every MMAs, like a smem-ring release;
// issuers = 2: a second warp issues an identical chain into a second accumulator at the same time
__global__ void __launch_bounds__(128, 1) probe_mma_kernel(int M, int N, int n_mma, int kblocks, int commit_every, int issuers, long long* clocks_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
  // A: kblocks tiles of [128 rows][128 B], B: kblocks tiles of [256 rows][128 B]
  const uint32_t a0 = base, b0 = base + (uint32_t)kblocks * 16384u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(al + (size_t)kblocks * (16384 + 32768));
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);  // bar+2: second issuer's barrier, bar+3: scratch
  for (int i = threadIdx.x; i < kblocks * (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(al)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_init(smem_u32(bar + 2), 1); mbar_init(smem_u32(bar + 3), 1u << 20); mbar_init_fence(); }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(slot));
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    const uint32_t idesc = idesc_bf16(M, N, false, false);
    const uint32_t mybar = smem_u32(bar + 2 * w), acc = tmem + 256u * w;
    const long long t0 = clock64();
    // kblocks == 4 and commit_every in {0, 4, 8, 16}: no divisions in the issue loop
    for (int i = 0; i < n_mma; i += 4) {
      const int kb = (i >> 2) & 3;
      const uint32_t a_lo = desc_lo_kmajor(a0 + kb * 16384), b_lo = desc_lo_kmajor(b0 + kb * 32768);
      umma_bf16_lo(acc, a_lo, b_lo, idesc, i > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 1; k < 4; k++) umma_bf16_lo(acc, a_lo + 2u * k, b_lo + 2u * k, idesc, 1u);
      if (commit_every > 0 && ((i + 4) & (commit_every - 1)) == 0) umma_commit(smem_u32(bar + 3));
    }
    umma_commit(mybar);
    const long long t1 = clock64();
    mbar_wait(mybar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && w == 0) { clocks_out[0] = t1 - t0; clocks_out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

bool probe_mma(cudaStream_t s, int M, int N, int n_mma, int commit_every, int issuers, long long* issue_clk, long long* total_clk) {
  const int kblocks = 4;
  const int smem = kblocks * (16384 + 32768) + 1024 + 64;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return false;
  probe_mma_kernel<<<148, 128, smem, s>>>(M, N, n_mma, kblocks, commit_every, issuers, d);
  long long hbuf[2] = {0, 0};
  const bool ok = cudaStreamSynchronize(s) == cudaSuccess && cudaMemcpy(hbuf, d, 16, cudaMemcpyDeviceToHost) == cudaSuccess;
  cudaFree(d);
  *issue_clk = hbuf[0]; *total_clk = hbuf[1];
  return ok;
}

}  // namespace lrcn
