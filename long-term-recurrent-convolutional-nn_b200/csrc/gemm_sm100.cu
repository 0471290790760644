// gemm_sm100.cu -- persistent tcgen05 / TMEM / TMA GEMM for sm_100a with error-compensated bf16 splitting.
//
//   C[M][N] (fp32) = Aop * Bop,  A = A_hi + A_lo, B = B_hi + B_lo (bf16 pairs, |lo| <= 2^-9 |hi|)
//   D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        (3 tcgen05.mma per 16-wide k-step, fp32 TMEM accumulate)
//
// which keeps the result within ~2^-16 relative of an fp32 SGEMM (measured 4.5e-6): the 1e-4 parity bar of the
// LRCN hot path cannot be met by single-pass bf16 (2.3e-3) or tf32 (2.9e-4) (SURVEY.md §7 hard part 1).
//
// Operands are row-major 2-D bf16 tensors, each K-major ([MN][K]: forward GEMMs, the reference's column-major
// K x N weights as they lie in memory) or MN-major ([K][MN]: the data-/weight-gradient GEMMs of BPTT), selected
// by the UMMA instruction-descriptor major bits -- no transposed copies are ever made.
//
// Persistent kernel, one CTA per SM, 192 threads:
//   warp 0      TMA producer   (smem ring of {A_hi,A_lo,B_hi,B_lo} stages, SWIZZLE_128B, mbarrier expect_tx)
//   warp 1      TMEM allocator + MMA issuer (one elected lane; tcgen05.commit frees stages / publishes accumulators)
//   warps 2-5   epilogue: tcgen05.ld (32x32b.x32) -> per-warp smem transpose -> coalesced float4 global stores
//               (+bias, +beta*C, optional bf16 hi/lo split of the result, or fp32 atomics under split-K)
// Two TMEM accumulator stages, so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tile 128 x BN x 64 with BN = 128 (3 smem stages) or 256 (2 stages).
#include "kernels.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>

namespace lrcn {

using namespace ptx;

static thread_local std::string g_gemm_err;
const char* gemm_bf16x3_last_error() { return g_gemm_err.c_str(); }
void set_sm100_error(const char* msg) { g_gemm_err = msg; }

constexpr int BM = 128, BK = 64;
constexpr int A_TILE = BM * BK * 2;  // 16 KiB
constexpr int NUM_THREADS = 192;
constexpr int EPI_LD = 36;                            // padded row (floats) of the per-warp transpose buffer (LSU epilogue)
constexpr int EPI_WARP_BYTES = 8192;                  // per epilogue warp: two 4 KiB TMA-store staging buffers (or one transpose buffer)
constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;

template <int BN_>
struct TileCfg {
  static constexpr int STAGES = BN_ == 128 ? 3 : 2;
  static constexpr int B_TILE = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN_;
};

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split;
  float* C; int ldc;
  const float* bias;
  int beta;
  __nv_bfloat16* C_hi; __nv_bfloat16* C_lo;
  int tma_epi;  // fp32 result leaves through TMA stores (TMA reduce-adds under split-K)
};

// CM = cluster size along M: the CM CTAs of a cluster compute M-adjacent tiles of the same N-tile in lockstep; each loads
// 1/CM of the B tile and multicasts it to the others, cutting the L2->SMEM traffic these GEMMs are bound by.
template <bool AK, bool BKM, int BN_, int CM>
__global__ void __cluster_dims__(CM, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = TileCfg<BN_>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi_buf = reinterpret_cast<float*>(smem_al + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * Cfg::STAGE_BYTES + EPI_BYTES);
  const uint32_t full_bar0 = smem_u32(bars);
  const uint32_t empty_bar0 = smem_u32(bars + STAGES);
  const uint32_t tfull_bar0 = smem_u32(bars + 2 * STAGES);      // 2 accumulator stages
  const uint32_t tempty_bar0 = smem_u32(bars + 2 * STAGES + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int tiles_mg = (p.tiles_m + CM - 1) / CM;  // groups of CM M-adjacent tiles
  const int tiles_mn = tiles_mg * p.tiles_n;
  const int total_work = tiles_mn * p.splits;
  const int rank = CM > 1 ? (int)cluster_ctarank() : 0;
  const int cidx = blockIdx.x / CM, ncl = gridDim.x / CM;
  constexpr uint16_t CMASK = (uint16_t)((1u << CM) - 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, CM); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar0 + 8 * a, 1); mbar_init(tempty_bar0 + 8 * a, 4); }
    mbar_init_fence();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA_hi); prefetch_tensormap(&tmA_lo); prefetch_tensormap(&tmB_hi); prefetch_tensormap(&tmB_lo);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (CM > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // everything above overlapped the previous kernel's tail; its results are visible from here on
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int w = cidx; w < total_work; w += ncl) {
        const int z = w / tiles_mn, rem = w - z * tiles_mn;
        const int m0 = ((rem % tiles_mg) * CM + rank) * BM, n0 = (rem / tiles_mg) * BN_;
        const int kb_begin = z * p.kb_per_split, kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar0 + 8 * s, ph ^ 1);  // every CTA of the cluster has consumed this stage
          const uint32_t full = full_bar0 + 8 * s;
          mbar_expect_tx(full, Cfg::STAGE_BYTES);
          const uint32_t sA_hi = smem_base + s * Cfg::STAGE_BYTES, sA_lo = sA_hi + A_TILE;
          const uint32_t sB_hi = sA_lo + A_TILE, sB_lo = sB_hi + Cfg::B_TILE;
          const int k0 = kb * BK;
          if (AK) {  // A tile [128 m][64 k], inner = k
            tma_load_2d(sA_hi, &tmA_hi, full, k0, m0);
            tma_load_2d(sA_lo, &tmA_lo, full, k0, m0);
          } else {   // two boxes [64 k][64 m], inner = m
#pragma unroll
            for (int b = 0; b < 2; b++) {
              tma_load_2d(sA_hi + b * 8192, &tmA_hi, full, m0 + 64 * b, k0);
              tma_load_2d(sA_lo + b * 8192, &tmA_lo, full, m0 + 64 * b, k0);
            }
          }
          if (CM == 1) {
            if (BKM) {
              tma_load_2d(sB_hi, &tmB_hi, full, k0, n0);
              tma_load_2d(sB_lo, &tmB_lo, full, k0, n0);
            } else {
#pragma unroll
              for (int b = 0; b < BN_ / 64; b++) {
                tma_load_2d(sB_hi + b * 8192, &tmB_hi, full, n0 + 64 * b, k0);
                tma_load_2d(sB_lo + b * 8192, &tmB_lo, full, n0 + 64 * b, k0);
              }
            }
          } else {  // this CTA fetches its 1/CM share of the B tile and multicasts it to the whole cluster
            if (BKM) {
              constexpr int ROWS = BN_ / CM;  // tensor map box = {64, ROWS}
              tma_load_2d_mcast(sB_hi + rank * ROWS * 128, &tmB_hi, full, k0, n0 + rank * ROWS, CMASK);
              tma_load_2d_mcast(sB_lo + rank * ROWS * 128, &tmB_lo, full, k0, n0 + rank * ROWS, CMASK);
            } else {
              constexpr int NB = BN_ / 64 / CM;  // 64-wide boxes per CTA
#pragma unroll
              for (int bb = 0; bb < NB; bb++) {
                const int b = rank * NB + bb;
                tma_load_2d_mcast(sB_hi + b * 8192, &tmB_hi, full, n0 + 64 * b, k0, CMASK);
                tma_load_2d_mcast(sB_lo + b * 8192, &tmB_lo, full, n0 + 64 * b, k0, CMASK);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BM, BN_, !AK, !BKM);
      int it = 0, local = 0;
      for (int w = cidx; w < total_work; w += ncl, local++) {
        const int z = w / tiles_mn;
        const int kb_begin = z * p.kb_per_split, kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
        const int acc = local & 1;
        const uint32_t aph = (local >> 1) & 1;
        mbar_wait(tempty_bar0 + 8 * acc, aph ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN_);
        for (int kb = kb_begin; kb < kb_end; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sA_hi = smem_base + s * Cfg::STAGE_BYTES, sA_lo = sA_hi + A_TILE;
          const uint32_t sB_hi = sA_lo + A_TILE, sB_lo = sB_hi + Cfg::B_TILE;
#pragma unroll
          for (int k = 0; k < BK / 16; k++) {
            const uint64_t a_hi = AK ? desc_kmajor(sA_hi, k) : desc_mnmajor(sA_hi, k);
            const uint64_t a_lo = AK ? desc_kmajor(sA_lo, k) : desc_mnmajor(sA_lo, k);
            const uint64_t b_hi = BKM ? desc_kmajor(sB_hi, k) : desc_mnmajor(sB_hi, k);
            const uint64_t b_lo = BKM ? desc_kmajor(sB_lo, k) : desc_mnmajor(sB_lo, k);
            umma_bf16(tmem_d, a_lo, b_hi, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);  // small terms first
            umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
            umma_bf16(tmem_d, a_hi, b_hi, idesc, 1u);
          }
          if (CM > 1) umma_commit_mcast(empty_bar0 + 8 * s, CMASK);  // stage consumed in MY smem: tell every producer of the cluster
          else umma_commit(empty_bar0 + 8 * s);                      // frees the smem stage once the MMAs above retire
        }
        umma_commit(tfull_bar0 + 8 * acc);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====================
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    float* tb = epi_buf + quad * (EPI_WARP_BYTES / 4);
    const uint32_t stage_buf = smem_u32(tb);
    const bool split = p.splits > 1;
    if (p.tma_epi && lane == 0) prefetch_tensormap(&tmC);
    uint32_t chunk_no = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    int local = 0;
    for (int w = cidx; w < total_work; w += ncl, local++) {
      const int z = w / tiles_mn, rem = w - z * tiles_mn;
      const int m0 = ((rem % tiles_mg) * CM + rank) * BM, n0 = (rem / tiles_mg) * BN_;
      const int acc = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      const bool use_bias = p.bias && (!split || z == 0);
      float bias_next = (use_bias && n0 + lane < p.N) ? p.bias[n0 + lane] : 0.f;
      mbar_wait(tfull_bar0 + 8 * acc, aph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN_; c += 32) {
        uint32_t v[32];
        LRCN_TMEM_LD_32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN_ + c), v);
        const float bias_lane = bias_next;
        if (p.tma_epi && use_bias && c + 32 < BN_ && n0 + c + 32 + lane < p.N) bias_next = p.bias[n0 + c + 32 + lane];
        tmem_ld_wait();
        if (c + 32 >= BN_) {  // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar0 + 8 * acc);
        }
        const int nb = n0 + c;
        if (nb >= p.N) continue;  // warp-uniform
        if (p.tma_epi) {
          if (m0 + quad * 32 < p.M)
            epilogue_chunk_tma(&tmC, stage_buf + (chunk_no & 1u) * 4096u, v, bias_lane, use_bias, nb, m0 + quad * 32, split || p.beta, lane);
          chunk_no++;
          continue;
        }
        // transpose through smem: lane i holds row i of the 32x32 chunk
#pragma unroll
        for (int q = 0; q < 8; q++)
          *reinterpret_cast<float4*>(tb + lane * EPI_LD + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        const int mrow0 = m0 + quad * 32;
        if (vec_ok && !split && nb + 32 <= p.N) {
          const int cq = lane & 7, rsub = lane >> 3;  // lane -> 4 columns [4cq,4cq+4) of row (4*it + rsub)
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bv = *reinterpret_cast<const float4*>(p.bias + nb + 4 * cq);
#pragma unroll
          for (int itr = 0; itr < 8; itr++) {
            const int rr = 4 * itr + rsub, m = mrow0 + rr;
            if (m < p.M) {
              float4 x = *reinterpret_cast<const float4*>(tb + rr * EPI_LD + 4 * cq);
              x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
              float* cp = p.C + (size_t)m * p.ldc + nb + 4 * cq;
              if (p.beta) { const float4 o = *reinterpret_cast<const float4*>(cp); x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
              *reinterpret_cast<float4*>(cp) = x;
              if (p.C_hi) {
                __nv_bfloat16 hh[4], ll[4];
                split_bf16(x.x, hh[0], ll[0]); split_bf16(x.y, hh[1], ll[1]); split_bf16(x.z, hh[2], ll[2]); split_bf16(x.w, hh[3], ll[3]);
                *reinterpret_cast<uint2*>(p.C_hi + (size_t)m * p.ldc + nb + 4 * cq) = *reinterpret_cast<uint2*>(hh);
                *reinterpret_cast<uint2*>(p.C_lo + (size_t)m * p.ldc + nb + 4 * cq) = *reinterpret_cast<uint2*>(ll);
              }
            }
          }
        } else {
          const int n = nb + lane;  // lane -> column
          const bool nok = n < p.N;
          const float bv = (p.bias && nok && (!split || z == 0)) ? p.bias[n] : 0.f;
#pragma unroll 4
          for (int rr = 0; rr < 32; rr++) {
            const int m = mrow0 + rr;
            if (m < p.M && nok) {
              float x = tb[rr * EPI_LD + lane] + bv;
              float* cp = p.C + (size_t)m * p.ldc + n;
              if (split) {
                atomicAdd(cp, x);  // C pre-zeroed by the launcher when !beta
              } else {
                if (p.beta) x += *cp;
                *cp = x;
                if (p.C_hi) {
                  __nv_bfloat16 hh, ll;
                  split_bf16(x, hh, ll);
                  p.C_hi[(size_t)m * p.ldc + n] = hh;
                  p.C_lo[(size_t)m * p.ldc + n] = ll;
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
    if (p.tma_epi && lane == 0) bulk_wait<0>();  // all of this warp's stores have been written before the CTA may exit
  }
  tc_fence_before();
  __syncthreads();
  if (CM > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into its smem / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
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
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu box_outer=%u", (int)r, (unsigned long long)inner,
             (unsigned long long)outer, (unsigned long long)ld, box_outer);
    g_gemm_err = buf;
    return false;
  }
  return true;
}

static bool make_map_f32_out(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { g_gemm_err = "cuTensorMapEncodeTiled unavailable"; return false; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 3)) { g_gemm_err = "fp32 output not 16-byte aligned / ld not a multiple of 4"; return false; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(f32 out) failed (%d) inner=%llu outer=%llu ld=%llu", (int)r, (unsigned long long)inner,
             (unsigned long long)outer, (unsigned long long)ld);
    g_gemm_err = buf;
    return false;
  }
  return true;
}

struct MapKey {
  const void* p; uint64_t inner, outer, ld; uint32_t box;  // box == 0xF32 marks the fp32 output maps
  bool operator<(const MapKey& o) const { return std::tie(p, inner, outer, ld, box) < std::tie(o.p, o.inner, o.outer, o.ld, o.box); }
};
static std::map<MapKey, CUtensorMap>& map_cache() { static std::map<MapKey, CUtensorMap> c; return c; }
static std::mutex g_map_mu;

bool get_tensor_map_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  MapKey k{ptr, inner, outer, ld, box_outer};
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto& c = map_cache();
  auto it = c.find(k);
  if (it != c.end()) { *out = it->second; return true; }
  CUtensorMap m;
  if (!make_map(&m, ptr, inner, outer, ld, box_outer)) return false;
  if (c.size() > 8192) c.clear();
  c[k] = m;
  *out = m;
  return true;
}

bool get_tensor_map_f32_out(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld) {
  MapKey k{ptr, inner, outer, ld, 0xF32u};
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto& c = map_cache();
  auto it = c.find(k);
  if (it != c.end()) { *out = it->second; return true; }
  CUtensorMap m;
  if (!make_map_f32_out(&m, ptr, inner, outer, ld)) return false;
  if (c.size() > 8192) c.clear();
  c[k] = m;
  *out = m;
  return true;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int g_num_sms = 148, g_force_bn = 0, g_force_cm = 0, g_two_cta = 1;
int g_gemm_tma_epi = 1;
int g_pdl = 1;  // measured (profiles/r01_progress.md): PDL on the tcgen05 kernels gains 1.3 %, adding the small SIMT kernels loses 2 %  // LRCN_GEMM_TMA_EPI=0 keeps the LSU epilogue everywhere

bool gemm_tma_epilogue_ok(const float* C, int ldc, bool beta, const void* C_hi) {
  (void)beta;  // beta = accumulate into C = TMA reduce-add
  return g_gemm_tma_epi && !C_hi && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
}

template <bool AK, bool BKM, int BN_>
static cudaError_t set_attr() {
  cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel<AK, BKM, BN_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<BN_>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16x3_kernel<AK, BKM, BN_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<BN_>::SMEM_BYTES);
  return e;
}
bool init_gemm_sm100() {
  cudaError_t e = set_attr<true, true, 128>();
  if (e == cudaSuccess) e = set_attr<true, false, 128>();
  if (e == cudaSuccess) e = set_attr<false, true, 128>();
  if (e == cudaSuccess) e = set_attr<false, false, 128>();
  if (e == cudaSuccess) e = set_attr<true, true, 256>();
  if (e == cudaSuccess) e = set_attr<true, false, 256>();
  if (e == cudaSuccess) e = set_attr<false, true, 256>();
  if (e == cudaSuccess) e = set_attr<false, false, 256>();
  if (e != cudaSuccess) { g_gemm_err = std::string("cudaFuncSetAttribute(gemm_bf16x3): ") + cudaGetErrorString(e); return false; }
  if (!get_encode()) { g_gemm_err = "cuTensorMapEncodeTiled unavailable"; return false; }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  g_force_bn = env_int("LRCN_GEMM_BN", 0);
  g_force_cm = env_int("LRCN_GEMM_CM", 0);
  g_two_cta = env_int("LRCN_GEMM_2CTA", 1);
  g_gemm_tma_epi = env_int("LRCN_GEMM_TMA_EPI", 1);
  g_pdl = env_int("LRCN_PDL", 1);
  if (g_two_cta && !init_gemm2_sm100()) return false;
  return true;
}

template <bool AK, bool BKM, int BN_>
static void launch(cudaStream_t s, int grid, int cm, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                   const CUtensorMap& b_lo, const CUtensorMap& c, const GemmParams& p) {
  if (cm == 2) launch_pdl(gemm_bf16x3_kernel<AK, BKM, BN_, 2>, dim3(grid), dim3(NUM_THREADS), TileCfg<BN_>::SMEM_BYTES, s, a_hi, a_lo, b_hi, b_lo, c, p);
  else launch_pdl(gemm_bf16x3_kernel<AK, BKM, BN_, 1>, dim3(grid), dim3(NUM_THREADS), TileCfg<BN_>::SMEM_BYTES, s, a_hi, a_lo, b_hi, b_lo, c, p);
}

bool gemm_bf16x3(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N, int K, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
                 int lda, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, float* C, int ldc, bool beta, const float* bias,
                 __nv_bfloat16* C_hi, __nv_bfloat16* C_lo, bool c_zeroed) {
  if (M <= 0 || N <= 0 || K <= 0) return true;
  const int tm = (M + BM - 1) / BM;
  // CTA-pair kernel (256 x 256 UMMA, half the B bytes per SM) whenever the output is wide enough to fill the pair tiles
  const int pair_tiles = ((tm + 1) / 2) * ((N + 255) / 256), pairs = g_num_sms / 2;
  if (g_two_cta && g_force_bn == 0 && tm >= 2 && N >= 192 &&
      (pair_tiles * 2 >= pairs || (!C_hi && (long long)pair_tiles * ((K + BK - 1) / BK) >= 6ll * pairs)))
    return gemm2_bf16x3(s, a_kmajor, b_kmajor, M, N, K, A_hi, A_lo, lda, B_hi, B_lo, ldb, C, ldc, beta, bias, C_hi, C_lo, c_zeroed);
  // BN = 256 halves the smem operand traffic per MMA (128x128 SS-mode MMAs sit right at the 128 B/clk smem limit);
  // use it when there is enough N to keep every SM busy
  int bn = (N >= 512 && tm * ((N + 255) / 256) >= g_num_sms) ? 256 : 128;
  if (g_force_bn == 128 || g_force_bn == 256) bn = g_force_bn;
  const int tn = (N + bn - 1) / bn;
  const int num_kb = (K + BK - 1) / BK;
  int splits = 1;
  if (!C_hi) {
    const int tiles = tm * tn;
    if (tiles * 2 <= g_num_sms && num_kb >= 8) {
      splits = g_num_sms / tiles;
      if (splits > num_kb / 4) splits = num_kb / 4;
      if (splits < 1) splits = 1;
      if (splits > 32) splits = 32;
    }
  }
  const int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;
  // optional cluster of 2 M-adjacent tiles sharing (multicasting) the B tile
  // Measured on B200 (profiles/r01_progress.md): the multicast does NOT speed these GEMMs up -- they are bound by the bytes each
  // SM can keep in flight (smem capacity / L2 latency), not by L2 read bandwidth -- so it is off unless LRCN_GEMM_CM=2 asks for it.
  int cm = 1;
  if (g_force_cm == 2 && tm >= 2) cm = 2;

  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  bool ok = true;
  if (a_kmajor) ok = ok && get_tensor_map_bf16(&ta_hi, A_hi, K, M, lda, BM) && get_tensor_map_bf16(&ta_lo, A_lo, K, M, lda, BM);
  else          ok = ok && get_tensor_map_bf16(&ta_hi, A_hi, M, K, lda, BK) && get_tensor_map_bf16(&ta_lo, A_lo, M, K, lda, BK);
  if (b_kmajor) ok = ok && get_tensor_map_bf16(&tb_hi, B_hi, K, N, ldb, bn / cm) && get_tensor_map_bf16(&tb_lo, B_lo, K, N, ldb, bn / cm);
  else          ok = ok && get_tensor_map_bf16(&tb_hi, B_hi, N, K, ldb, BK) && get_tensor_map_bf16(&tb_lo, B_lo, N, K, ldb, BK);
  if (!ok) return false;
  const bool tma_epi = gemm_tma_epilogue_ok(C, ldc, beta, C_hi);
  CUtensorMap tc = ta_hi;  // placeholder when unused
  if (tma_epi && !get_tensor_map_f32_out(&tc, C, N, M, ldc)) return false;

  if (splits > 1 && !beta && !c_zeroed) {
    if (ldc == N) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), s);
    else cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.tiles_m = tm; p.tiles_n = tn; p.splits = splits; p.kb_per_split = kb_per;
  p.C = C; p.ldc = ldc; p.bias = bias; p.beta = beta ? 1 : 0; p.C_hi = C_hi; p.C_lo = C_lo; p.tma_epi = tma_epi ? 1 : 0;
  const int total = ((tm + cm - 1) / cm) * tn * splits;  // work units per cluster
  const int max_cl = g_num_sms / cm;
  const int grid = (total < max_cl ? total : max_cl) * cm;
#define LRCN_LAUNCH(AKv, BKv)                                                         \
  do {                                                                                \
    if (bn == 256) launch<AKv, BKv, 256>(s, grid, cm, ta_hi, ta_lo, tb_hi, tb_lo, tc, p); \
    else launch<AKv, BKv, 128>(s, grid, cm, ta_hi, ta_lo, tb_hi, tb_lo, tc, p);           \
  } while (0)
  if (a_kmajor && b_kmajor) LRCN_LAUNCH(true, true);
  else if (a_kmajor && !b_kmajor) LRCN_LAUNCH(true, false);
  else if (!a_kmajor && b_kmajor) LRCN_LAUNCH(false, true);
  else LRCN_LAUNCH(false, false);
#undef LRCN_LAUNCH
  if (g_counter) g_counter->n++;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { g_gemm_err = std::string("gemm_bf16x3 launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

bool gemm_bind_abort(unsigned int* host_flag) { return dev_abort_bind(host_flag) == cudaSuccess; }

}  // namespace lrcn
