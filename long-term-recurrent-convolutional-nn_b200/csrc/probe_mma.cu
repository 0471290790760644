// probe_mma.cu -- diagnostics: issue-to-completion cost of a chain of tcgen05.mma (kind::f16, bf16, SS mode, K-major
// SWIZZLE_128B operands resident in smem) accumulating into one TMEM tile, for a given M x N.  One CTA per SM.
#include "kernels.cuh"
#include "sm100_ptx.cuh"

namespace lrcn {
using namespace ptx;

// commit_every > 0: a tcgen05.commit (to a scratch mbarrier with a huge count) after every commit_every MMAs, like a smem-ring release;
// issuers = 2: a second warp issues an identical chain into a second accumulator at the same time
// issuers >= 16: flag bits on top of the issuer count (low 4 bits): 16 = A operand from TMEM (TS form), 32 / 64 / 128 = 4 / 8 / 12 extra warps
// spinning on an mbarrier try_wait (what the waiting epilogue warps of the LSTM kernels do), 256 = one warp polling global memory
__device__ __forceinline__ void umma_bf16_ts_probe(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI_SW128)
      : "memory");
}
__global__ void __launch_bounds__(1024, 1) probe_mma_kernel(int M, int N, int n_mma, int kblocks, int commit_every, int issuers_flags, long long* clocks_out, unsigned int* gpoll) {
  const int issuers = issuers_flags & 15;
  const bool ts = issuers_flags & 16;
  __shared__ volatile int done_flag;
  if (threadIdx.x == 0) done_flag = 0;
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
  if (w >= 4) {  // background warps
    if (w == 4 && (issuers_flags & 256)) {
      if ((threadIdx.x & 31) == 0) while (!done_flag) { unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gpoll) : "memory"); if (v == 0xdeadbeefu) break; }
    } else {
      while (!done_flag) { if (mbar_try_wait(smem_u32(bar + 3), 0)) break; }
    }
  }
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    const uint32_t idesc = idesc_bf16(M, N, false, false);
    const uint32_t mybar = smem_u32(bar + 2 * w), acc = tmem + 256u * w;
    const long long t0 = clock64();
    // kblocks == 4 and commit_every in {0, 4, 8, 16}: no divisions in the issue loop
    for (int i = 0; i < n_mma; i += 4) {
      const int kb = (i >> 2) & 3;
      const uint32_t a_lo = desc_lo_kmajor(a0 + kb * 16384), b_lo = desc_lo_kmajor(b0 + kb * 32768);
      if (ts) {
        const uint32_t a_col = tmem + 384u + 32u * kb;  // any columns outside the two accumulators: the values do not matter
#pragma unroll
        for (int k = 0; k < 4; k++) umma_bf16_ts_probe(acc, a_col + 8u * k, b_lo + 2u * k, idesc, (i | k) ? 1u : 0u);
      } else {
        umma_bf16_lo(acc, a_lo, b_lo, idesc, i > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < 4; k++) umma_bf16_lo(acc, a_lo + 2u * k, b_lo + 2u * k, idesc, 1u);
      }
      if (commit_every > 0 && ((i + 4) & (commit_every - 1)) == 0) umma_commit(smem_u32(bar + 3));
    }
    umma_commit(mybar);
    const long long t1 = clock64();
    mbar_wait(mybar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && w == 0) { clocks_out[0] = t1 - t0; clocks_out[1] = t2 - t0; }
    if (w == 0) done_flag = 1;
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
  if (cudaMalloc(&d, 32) != cudaSuccess) return false;
  cudaMemsetAsync(d, 0, 32, s);
  const int extra = ((issuers & 32) ? 4 : 0) + ((issuers & 64) ? 8 : 0) + ((issuers & 128) ? 12 : 0) + ((issuers & 256) ? 1 : 0);
  probe_mma_kernel<<<148, 128 + 32 * extra, smem, s>>>(M, N, n_mma, kblocks, commit_every, issuers, d, reinterpret_cast<unsigned int*>(d + 2));
  long long hbuf[2] = {0, 0};
  const bool ok = cudaStreamSynchronize(s) == cudaSuccess && cudaMemcpy(hbuf, d, 16, cudaMemcpyDeviceToHost) == cudaSuccess;
  cudaFree(d);
  *issue_clk = hbuf[0]; *total_clk = hbuf[1];
  return ok;
}

}  // namespace lrcn
