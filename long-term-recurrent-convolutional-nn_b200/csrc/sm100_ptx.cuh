// sm100_ptx.cuh -- inline-PTX building blocks for sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld), UMMA shared-memory and instruction descriptors.
// Encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) and were validated on
// B200 by tools/probe_gemm.py for K-major and MN-major bf16 operands with SWIZZLE_128B.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "kernels.cuh"

namespace lrcn {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// bounded wait: a protocol bug must end the kernel (flag + drain, kernels.cuh) instead of hanging the GPU box
constexpr long long WAIT_LIMIT_CLK = 4000000000ll;  // ~2 s
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0u && dev_aborted()) return;
    if (clock64() - t0 > WAIT_LIMIT_CLK) {
      printf("lrcn sm100: mbarrier timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      dev_abort_set();
      return;
    }
  }
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// smem matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major SW128 tile ([rows][64 bf16], 8-row groups of 1024 B): advance 16 k-elements = 32 B; SBO = 1024
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int k16) { return smem_desc(tile + (uint32_t)k16 * 32u, 16u, 1024u); }
// MN-major SW128 tile (boxes of [64 k-rows][64 mn] = 8192 B each): advance 16 k-rows = 2048 B; LBO = 8192 (next 64-wide
// MN chunk), SBO = 1024 (next 8 k-rows)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int k16) { return smem_desc(tile + (uint32_t)k16 * 2048u, 8192u, 1024u); }
// instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, a_major bit15, b_major bit16
// (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Cheap descriptor arithmetic for MMA issue loops.  One thread issues every tcgen05.mma; rebuilding the 64-bit descriptors
// with shifts and masks costs ~14 dependent uniform-datapath instructions (~85-100 clk, measured with tools/probe_mma_rate.py)
// per MMA -- more than a 128x128x16 MMA takes on the tensor pipe (64 clk).  The high word of a descriptor is constant per
// layout; the low word is (addr >> 4) | (LBO >> 4) << 16, so advancing k is ONE 32-bit add.
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_kmajor(uint32_t tile) { return ((tile & 0x3FFFFu) >> 4) | (1u << 16); }   // +2 per k16 step (32 B)
__device__ __forceinline__ uint32_t desc_lo_mnmajor(uint32_t tile) { return ((tile & 0x3FFFFu) >> 4) | (512u << 16); }  // +128 per k16 step (2048 B)
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI_SW128)
      : "memory");
}

// One lane of a CONVERGED warp.  tcgen05.mma takes its descriptors from uniform registers: inside `if (lane == 0)` the compiler
// cannot prove that a single thread is active and wraps every MMA in an ELECT / R2UR x4 / branch "waterfall" (~125 clk per MMA
// in the LSTM issue loops); warp-converged loops over warp-uniform values + elect.sync let it keep the descriptors in uniform
// registers and advance them with uniform-datapath adds.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

// ---- TMA stores (smem -> global), bulk async-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#define LRCN_TMEM_LD_32(taddr, v)                                                                                  \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18," \
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                               \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
      : "r"(taddr)                                                                                                 \
      : "memory")
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}  // namespace ptx

// host: cached 2-D bf16 tensor maps (row-major [outer][inner], pitch ld elements, box {64, box_outer}, SWIZZLE_128B)
bool get_tensor_map_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer);
// fp32 output map for the TMA-store epilogue: row-major [outer][inner], pitch ld floats (multiple of 4), box {32, 32}, SWIZZLE_128B
bool get_tensor_map_f32_out(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld);

// Epilogue of one 32-row x 32-column accumulator chunk through a TMA store (or TMA reduce-add for split-K partial sums):
// lane r holds row r (v[0..31]); `buf` is this warp's 4 KiB, 1024 B-aligned staging buffer for this chunk parity.
// At most one older bulk group of this thread may still be reading the other buffer.
__device__ __forceinline__ void epilogue_chunk_tma(const CUtensorMap* tmC, uint32_t buf, const uint32_t (&v)[32], float bias_lane, bool has_bias,
                                                   int col0, int row0, bool reduce, int lane) {
  if (lane == 0) ptx::bulk_wait_read<1>();
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; q++) {
    float x0 = __uint_as_float(v[4 * q]), x1 = __uint_as_float(v[4 * q + 1]), x2 = __uint_as_float(v[4 * q + 2]), x3 = __uint_as_float(v[4 * q + 3]);
    if (has_bias) {
      x0 += __shfl_sync(0xffffffffu, bias_lane, 4 * q);
      x1 += __shfl_sync(0xffffffffu, bias_lane, 4 * q + 1);
      x2 += __shfl_sync(0xffffffffu, bias_lane, 4 * q + 2);
      x3 += __shfl_sync(0xffffffffu, bias_lane, 4 * q + 3);
    }
    // SWIZZLE_128B: 16-byte chunk q of row r lives at chunk (q ^ (r & 7)) -> the 8 lanes of a store phase hit 8 distinct bank groups
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4)), "f"(x0), "f"(x1),
                 "f"(x2), "f"(x3)
                 : "memory");
  }
  ptx::fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (reduce) ptx::tma_reduce_add_2d(tmC, buf, col0, row0);
    else ptx::tma_store_2d(tmC, buf, col0, row0);
    ptx::bulk_commit();
  }
}
void set_sm100_error(const char* msg);

}  // namespace lrcn
