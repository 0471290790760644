// gemm_sm100.cu -- tcgen05 / TMEM / TMA GEMM for sm_100a with error-compensated bf16 splitting.
//
//   C[M][N] (fp32) = Aop * Bop,  A = A_hi + A_lo, B = B_hi + B_lo (bf16 pairs, |lo| <= 2^-9 |hi|)
//   D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (3 tcgen05.mma per 16-wide k-step, fp32 TMEM accumulate)
//
// which keeps the result within ~2^-16 relative of an fp32 SGEMM: the 1e-4 parity bar of the LRCN
// hot path (SURVEY.md §7 "hard parts" 1) cannot be met by single-pass bf16 (2.3e-3) or tf32 (2.9e-4).
//
// Operands are row-major 2-D bf16 tensors; each may be K-major ([MN][K], forward GEMMs, the
// reference's column-major K x N weights) or MN-major ([K][MN], the data/weight-gradient GEMMs of BPTT),
// selected by the UMMA instruction-descriptor major bits, so no transposed copies are ever made.
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one elected lane),
// warps 2-5 epilogue (tcgen05.ld of the 4 TMEM lane quadrants).  3-stage smem ring
// (A_hi,A_lo,B_hi,B_lo tiles of 128x64 bf16, SWIZZLE_128B), mbarrier full/empty pipeline,
// tcgen05.commit releases stages and signals the epilogue.  Tile 128 x 128 x 64, optional split-K.
#include "kernels.cuh"
#include <cuda.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

namespace lrcn {

static thread_local std::string g_gemm_err;
const char* gemm_bf16x3_last_error() { return g_gemm_err.c_str(); }

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 2;          // 16 KiB per operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 128;

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// bounded wait: a protocol bug must trap (launch failure) instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) { printf("lrcn gemm_sm100: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
// smem matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

#define TMEM_LD_32(taddr, v)                                                                                      \
  asm volatile(                                                                                                   \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18," \
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                               \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
      : "r"(taddr)                                                                                                \
      : "memory")

struct GemmParams {
  int M, N, K;
  int kb_per_split;  // k-blocks (of BK) per grid.z slice
  float* C; int ldc;
  const float* bias;
  int beta;
  __nv_bfloat16* C_hi; __nv_bfloat16* C_lo;
  int mn_lbo, mn_sbo;  // MN-major descriptor strides (bytes)
};

template <bool AK, bool BKM>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE_BYTES);
  const uint32_t full_bar0 = smem_u32(bars);                 // STAGES barriers
  const uint32_t empty_bar0 = smem_u32(bars + STAGES);       // STAGES barriers
  const uint32_t tmem_full_bar = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < num_kb; i++) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(empty_bar0 + 8 * s, ph ^ 1);
        const uint32_t full = full_bar0 + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t st = smem_base + s * STAGE_BYTES;
        const int k0 = (kb_begin + i) * BK;
        if (AK) {  // A tile: [128 m][64 k], inner = k
          tma_load_2d(st, &tmA_hi, full, k0, m0);
          tma_load_2d(st + TILE_BYTES, &tmA_lo, full, k0, m0);
        } else {   // A tile: two boxes [64 k][64 m], inner = m
          tma_load_2d(st, &tmA_hi, full, m0, k0);
          tma_load_2d(st + TILE_BYTES / 2, &tmA_hi, full, m0 + 64, k0);
          tma_load_2d(st + TILE_BYTES, &tmA_lo, full, m0, k0);
          tma_load_2d(st + TILE_BYTES + TILE_BYTES / 2, &tmA_lo, full, m0 + 64, k0);
        }
        if (BKM) {
          tma_load_2d(st + 2 * TILE_BYTES, &tmB_hi, full, k0, n0);
          tma_load_2d(st + 3 * TILE_BYTES, &tmB_lo, full, k0, n0);
        } else {
          tma_load_2d(st + 2 * TILE_BYTES, &tmB_hi, full, n0, k0);
          tma_load_2d(st + 2 * TILE_BYTES + TILE_BYTES / 2, &tmB_hi, full, n0 + 64, k0);
          tma_load_2d(st + 3 * TILE_BYTES, &tmB_lo, full, n0, k0);
          tma_load_2d(st + 3 * TILE_BYTES + TILE_BYTES / 2, &tmB_lo, full, n0 + 64, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
      // a_major bit15, b_major bit16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((AK ? 0u : 1u) << 15) | ((BKM ? 0u : 1u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < num_kb; i++) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(full_bar0 + 8 * s, ph);
        tcgen05_fence_after();
        const uint32_t st = smem_base + s * STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; k++) {
          // K-major: 16 k-elements = 32 B inside the 128 B swizzle row; SBO = 1024 B (8 rows x 128 B)
          // MN-major: 16 k-rows = 2 x (8 rows x 128 B); LBO = stride between 64-wide MN chunks, SBO = 1024 B
          const uint32_t a_off = AK ? (uint32_t)k * 32u : (uint32_t)k * 2048u;
          const uint32_t b_off = BKM ? (uint32_t)k * 32u : (uint32_t)k * 2048u;
          const uint32_t a_lbo = AK ? 16u : (uint32_t)p.mn_lbo, a_sbo = AK ? 1024u : (uint32_t)p.mn_sbo;
          const uint32_t b_lbo = BKM ? 16u : (uint32_t)p.mn_lbo, b_sbo = BKM ? 1024u : (uint32_t)p.mn_sbo;
          const uint64_t a_hi = make_smem_desc(st + a_off, a_lbo, a_sbo);
          const uint64_t a_lo = make_smem_desc(st + TILE_BYTES + a_off, a_lbo, a_sbo);
          const uint64_t b_hi = make_smem_desc(st + 2 * TILE_BYTES + b_off, b_lbo, b_sbo);
          const uint64_t b_lo = make_smem_desc(st + 3 * TILE_BYTES + b_off, b_lbo, b_sbo);
          umma_bf16(tmem_base, a_lo, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);  // small terms first
          umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_bf16(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(empty_bar0 + 8 * s);  // frees the smem stage when the MMAs above retire
      }
      umma_commit(tmem_full_bar);         // accumulator complete -> epilogue
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int m = m0 + quad * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const bool split = gridDim.z > 1;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c;
      TMEM_LD_32(taddr, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (num_kb <= 0) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0u;
      }
      if (m < p.M) {
        const int nb = n0 + c;
        float* crow = p.C + (size_t)m * p.ldc + nb;
        if (split) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (nb + j < p.N) {
              float x = __uint_as_float(v[j]);
              if (p.bias && blockIdx.z == 0) x += p.bias[nb + j];
              atomicAdd(crow + j, x);
            }
          }
        } else if (vec_ok && nb + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 x = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (p.bias) { float4 b = *reinterpret_cast<const float4*>(p.bias + nb + j); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
            if (p.beta) { float4 o = *reinterpret_cast<const float4*>(crow + j); x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
            *reinterpret_cast<float4*>(crow + j) = x;
            if (p.C_hi) {
              __nv_bfloat16 h[4], l[4];
              const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int q = 0; q < 4; q++) { h[q] = __float2bfloat16_rn(xs[q]); l[q] = __float2bfloat16_rn(xs[q] - __bfloat162float(h[q])); }
              *reinterpret_cast<uint2*>(p.C_hi + (size_t)m * p.ldc + nb + j) = *reinterpret_cast<uint2*>(h);
              *reinterpret_cast<uint2*>(p.C_lo + (size_t)m * p.ldc + nb + j) = *reinterpret_cast<uint2*>(l);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (nb + j < p.N) {
              float x = __uint_as_float(v[j]);
              if (p.bias) x += p.bias[nb + j];
              if (p.beta) x += crow[j];
              crow[j] = x;
              if (p.C_hi) {
                __nv_bfloat16 h = __float2bfloat16_rn(x);
                p.C_hi[(size_t)m * p.ldc + nb + j] = h;
                p.C_lo[(size_t)m * p.ldc + nb + j] = __float2bfloat16_rn(x - __bfloat162float(h));
              }
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)f;
  }
  return fn;
}

// 2-D bf16 row-major tensor [outer][inner] with pitch ld (elements); box = {64 inner, box_outer}
static bool make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { g_gemm_err = "cuTensorMapEncodeTiled unavailable"; return false; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7)) { g_gemm_err = "bf16 operand not 16-byte aligned / ld not a multiple of 8"; return false; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r, (unsigned long long)inner,
             (unsigned long long)outer, (unsigned long long)ld);
    g_gemm_err = buf;
    return false;
  }
  return true;
}

struct MapKey {
  const void* p; uint64_t inner, outer, ld; uint32_t box;
  bool operator<(const MapKey& o) const { return std::tie(p, inner, outer, ld, box) < std::tie(o.p, o.inner, o.outer, o.ld, o.box); }
};
static std::map<MapKey, CUtensorMap>& map_cache() { static std::map<MapKey, CUtensorMap> c; return c; }
static std::mutex g_map_mu;

static bool get_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  MapKey k{ptr, inner, outer, ld, box_outer};
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto& c = map_cache();
  auto it = c.find(k);
  if (it != c.end()) { *out = it->second; return true; }
  CUtensorMap m;
  if (!make_map(&m, ptr, inner, outer, ld, box_outer)) return false;
  if (c.size() > 4096) c.clear();
  c[k] = m;
  *out = m;
  return true;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static int g_mn_lbo = 8192, g_mn_sbo = 1024;
bool init_gemm_sm100() {
  cudaError_t e = cudaSuccess;
  e = cudaFuncSetAttribute(gemm_bf16x3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { g_gemm_err = std::string("cudaFuncSetAttribute(gemm_bf16x3): ") + cudaGetErrorString(e); return false; }
  if (!get_encode()) { g_gemm_err = "cuTensorMapEncodeTiled unavailable"; return false; }
  g_mn_lbo = env_int("LRCN_MN_LBO", 8192);  // debug knobs for the MN-major descriptor strides
  g_mn_sbo = env_int("LRCN_MN_SBO", 1024);
  return true;
}

bool gemm_bf16x3(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N, int K, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
                 int lda, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, float* C, int ldc, bool beta, const float* bias,
                 __nv_bfloat16* C_hi, __nv_bfloat16* C_lo) {
  if (M <= 0 || N <= 0 || K <= 0) return true;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  bool ok = true;
  if (a_kmajor) { ok = ok && get_map(&ta_hi, A_hi, K, M, lda, BM) && get_map(&ta_lo, A_lo, K, M, lda, BM); }
  else          { ok = ok && get_map(&ta_hi, A_hi, M, K, lda, BK) && get_map(&ta_lo, A_lo, M, K, lda, BK); }
  if (b_kmajor) { ok = ok && get_map(&tb_hi, B_hi, K, N, ldb, BN) && get_map(&tb_lo, B_lo, K, N, ldb, BN); }
  else          { ok = ok && get_map(&tb_hi, B_hi, N, K, ldb, BK) && get_map(&tb_lo, B_lo, N, K, ldb, BK); }
  if (!ok) return false;

  const int tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  int splits = 1;
  if (!C_hi) {
    const int tiles = tm * tn;
    if (tiles < 96 && num_kb >= 8) {
      splits = (148 + tiles - 1) / tiles;
      if (splits > num_kb / 4) splits = num_kb / 4;
      if (splits < 1) splits = 1;
      if (splits > 32) splits = 32;
    }
  }
  int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;
  if (splits > 1 && !beta) {
    if (ldc == N) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), s);
    else cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.kb_per_split = kb_per; p.C = C; p.ldc = ldc; p.bias = bias; p.beta = beta ? 1 : 0;
  p.C_hi = C_hi; p.C_lo = C_lo;
  p.mn_lbo = g_mn_lbo;
  p.mn_sbo = g_mn_sbo;
  dim3 grid(tn, tm, splits);
  if (a_kmajor && b_kmajor) gemm_bf16x3_kernel<true, true><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  else if (a_kmajor && !b_kmajor) gemm_bf16x3_kernel<true, false><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  else if (!a_kmajor && b_kmajor) gemm_bf16x3_kernel<false, true><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  else gemm_bf16x3_kernel<false, false><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  if (g_counter) g_counter->n++;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_gemm_err = std::string("gemm_bf16x3 launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

}  // namespace lrcn
