// gemm2_sm100.cu -- 2-CTA (cta_group::2) variant of the bf16x3 tcgen05 GEMM.
//
// Why: the split-precision GEMMs move 4 B per operand element (bf16 hi + lo), and on B200 they are bound by the operand
// bytes each SM can pull through its shared memory (measured: ~8 TB/s chip-wide, profiles/r01_progress.md), not by the
// tensor pipe.  A CTA pair executing ONE 256 x 256 UMMA halves the B-operand traffic per SM: each CTA stages its own 128
// rows of A and only HALF (128 of 256 rows) of the B tile; the tensor cores of both SMs read both halves.
//   per CTA and k-block: A 2x16 KiB + B/2 2x16 KiB = 64 KiB for a 128 x 256 output slab (the 1-CTA kernel needs 96 KiB).
//
// Structure (per CTA, 192 threads; cluster = {leader, peer} along M):
//   warp 0  TMA producer: cp.async.bulk.tensor ... .cta_group::2, transaction bytes of BOTH CTAs land on the LEADER's
//           full barrier (mbarrier address with the peer bit cleared)
//   warp 1  TMEM allocator (cta_group::2, 512 columns = 2 accumulator stages x 256); in the leader: MMA issuer,
//           tcgen05.mma.cta_group::2, tcgen05.commit multicast to both CTAs' empty / accumulator-full barriers
//   warps 2-5 epilogue on the CTA's own 128 TMEM lanes; accumulator release = arrive on the leader's barrier
#include "kernels.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>

#include <string>

namespace lrcn {
using namespace ptx;

namespace g2 {
constexpr int BM = 128, BNP = 256, BNH = 128, BK = 64, STAGES = 3;
constexpr int TILE = 128 * BK * 2;               // 16 KiB: 128 rows x 128 B
constexpr int STAGE_BYTES = 4 * TILE;            // A_hi, A_lo, Bhalf_hi, Bhalf_lo
constexpr int EPI_LD = 36;
constexpr int EPI_WARP_BYTES = 8192;             // two 4 KiB TMA-store staging buffers per epilogue warp (or one transpose buffer)
constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // cute::Sm100MmaPeerBitMask: address the even (leader) CTA of the pair
constexpr int NUM_THREADS = 192;

__device__ __forceinline__ void tmem_alloc2(uint32_t slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(TMEM_COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// descriptors as (low word, constant high word): see desc_lo_kmajor / desc_lo_mnmajor in sm100_ptx.cuh
__device__ __forceinline__ void umma2_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI_SW128)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
}  // namespace g2

struct Gemm2Params {
  int M, N, K;
  int tiles_mp, tiles_n;  // tiles_mp: pairs of 128-row tiles
  int n_split;            // > 0: B is the column concatenation [B1 | B2]; tiles with n0 >= n_split read B2 (at n0 - n_split)
  int sk_tile0, sk_tiles, dp_end;  // SegIter: stream-K tiles [sk_tile0, +sk_tiles), whole tiles [0, dp_end)
  float* C; int ldc;
  const float* bias;
  int beta;
  __nv_bfloat16* C_hi; __nv_bfloat16* C_lo;
  int tma_epi;  // fp32 result leaves through TMA stores (TMA reduce-adds under split-K)
  int dbg;  // lrcn_bench_gemm diagnostics: 1 = no epilogue stores, 2 = no TMA loads after the first ring fill, 4 = no MMAs
};

// One unit of work of a CTA pair: k-blocks [kb0, kb1) of one output tile.  Stream-K ranges are cut at
// cidx * total / ncl and snapped to the tile boundary when they would leave a sliver of < 3 k-blocks.
struct Seg { int tile, kb0, kb1; };
// Work of one CTA pair: first its share of the stream-K tiles [sk_tile0, sk_tile0 + sk_tiles) -- an equal contiguous range
// of their (tile, k-block) iterations -- then whole tiles cidx, cidx + ncl, ... below dp_end.  Pure data-parallel: sk_tiles = 0;
// pure stream-K: dp_end = 0; hybrid (a last partial wave of tiles): both.
struct SegIter {
  int w, ncl, num_kb, it, it_end, sk_tile0, dp_end;
  __device__ __forceinline__ static int boundary(int c, int ncl, int tiles, int num_kb) {
    if (c >= ncl) return tiles * num_kb;
    int b = (int)(((long long)c * tiles * num_kb) / ncl);
    const int off = b % num_kb;
    if (off < 3) b -= off;
    else if (num_kb - off < 3) b += num_kb - off;
    return b;
  }
  __device__ __forceinline__ SegIter(int cidx, int ncl_, int num_kb_, int sk_tile0_, int sk_tiles, int dp_end_)
      : w(cidx), ncl(ncl_), num_kb(num_kb_), it(0), it_end(0), sk_tile0(sk_tile0_), dp_end(dp_end_) {
    if (sk_tiles > 0) { it = boundary(cidx, ncl, sk_tiles, num_kb); it_end = boundary(cidx + 1, ncl, sk_tiles, num_kb); }
  }
  __device__ __forceinline__ bool next(Seg& s) {
    if (it < it_end) {
      const int t = it / num_kb;
      s.tile = sk_tile0 + t; s.kb0 = it - t * num_kb;
      s.kb1 = min(num_kb, s.kb0 + (it_end - it));
      it += s.kb1 - s.kb0;
      return true;
    }
    if (w >= dp_end) return false;
    s.tile = w; s.kb0 = 0; s.kb1 = num_kb; w += ncl;
    return true;
  }
};

template <bool AK, bool BKM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::NUM_THREADS, 1)
gemm2_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const __grid_constant__ CUtensorMap tmB2_hi, const __grid_constant__ CUtensorMap tmB2_lo,
                    const __grid_constant__ CUtensorMap tmC, const Gemm2Params p) {
  using namespace g2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  float* epi_buf = reinterpret_cast<float*>(smem_al + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE_BYTES + EPI_BYTES);
  const uint32_t full_bar0 = smem_u32(bars);                 // used in the leader only
  const uint32_t empty_bar0 = smem_u32(bars + STAGES);       // one per CTA, signalled by the leader's multicast commit
  const uint32_t tfull_bar0 = smem_u32(bars + 2 * STAGES);   // one per CTA
  const uint32_t tempty_bar0 = smem_u32(bars + 2 * STAGES + 2);  // leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int tiles_mn = p.tiles_mp * p.tiles_n;
  const int cidx = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar0 + 8 * a, 1); mbar_init(tempty_bar0 + 8 * a, 8); }
    mbar_init_fence();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA_hi); prefetch_tensormap(&tmA_lo); prefetch_tensormap(&tmB_hi); prefetch_tensormap(&tmB_lo);
    if (p.n_split > 0) { prefetch_tensormap(&tmB2_hi); prefetch_tensormap(&tmB2_lo); }
  }
  if (warp == 1) tmem_alloc2(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int it = 0;
      SegIter si(cidx, ncl, num_kb_total, p.sk_tile0, p.sk_tiles, p.dp_end);
      Seg sg;
      while (si.next(sg)) {
        const int rem = sg.tile;
        const int m0 = ((rem % p.tiles_mp) * 2 + rank) * BM;
        int nh0 = (rem / p.tiles_mp) * BNP + rank * BNH;  // this CTA's half of the B tile
        const bool second = p.n_split > 0 && nh0 >= p.n_split;
        if (second) nh0 -= p.n_split;
        const CUtensorMap* mB_hi = second ? &tmB2_hi : &tmB_hi;
        const CUtensorMap* mB_lo = second ? &tmB2_lo : &tmB_lo;
        const int kb_begin = sg.kb0, kb_end = sg.kb1;
        for (int kb = kb_begin; kb < kb_end; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar0 + 8 * s, ph ^ 1);
          const uint32_t full = (full_bar0 + 8 * s) & PEER_MASK;  // the leader's barrier collects both CTAs' bytes
          if ((p.dbg & 2) && it >= STAGES) {
            if (leader) mbar_arrive(full_bar0 + 8 * s);
            continue;
          }
          if (leader) mbar_expect_tx(full_bar0 + 8 * s, 2 * STAGE_BYTES);
          const uint32_t sA_hi = smem_base + s * STAGE_BYTES, sA_lo = sA_hi + TILE, sB_hi = sA_lo + TILE, sB_lo = sB_hi + TILE;
          const int k0 = kb * BK;
          if (AK) {
            tma_load_2d_2sm(sA_hi, &tmA_hi, full, k0, m0);
            tma_load_2d_2sm(sA_lo, &tmA_lo, full, k0, m0);
          } else {
#pragma unroll
            for (int b = 0; b < 2; b++) {
              tma_load_2d_2sm(sA_hi + b * 8192, &tmA_hi, full, m0 + 64 * b, k0);
              tma_load_2d_2sm(sA_lo + b * 8192, &tmA_lo, full, m0 + 64 * b, k0);
            }
          }
          if (BKM) {
            tma_load_2d_2sm(sB_hi, mB_hi, full, k0, nh0);
            tma_load_2d_2sm(sB_lo, mB_lo, full, k0, nh0);
          } else {
#pragma unroll
            for (int b = 0; b < 2; b++) {
              tma_load_2d_2sm(sB_hi + b * 8192, mB_hi, full, nh0 + 64 * b, k0);
              tma_load_2d_2sm(sB_lo + b * 8192, mB_lo, full, nh0 + 64 * b, k0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the loop CONVERGED on warp-uniform values and one elected lane issues: the descriptors then live in
    // uniform registers and advance by one uniform add per k-step (inside `if (lane == 0)` every tcgen05.mma is wrapped in an
    // ELECT / R2UR / branch waterfall -- ~125 clk per instruction, as much as a 256x256x16 pair MMA takes on the tensor pipe).
    if (__shfl_sync(0xffffffffu, leader ? 1 : 0, 0)) {
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = idesc_bf16(256, BNP, !AK, !BKM);
      constexpr uint32_t A_STEP = AK ? 2u : 128u, B_STEP = BKM ? 2u : 128u;  // low-word advance per 16 k-elements (32 B K-major, 2048 B MN-major)
      int it = 0, local = 0;
      SegIter si(cidx, ncl, num_kb_total, p.sk_tile0, p.sk_tiles, p.dp_end);
      Seg sg;
      for (; si.next(sg); local++) {
        const int kb_begin = sg.kb0, kb_end = sg.kb1;
        const int acc = local & 1;
        const uint32_t aph = (local >> 1) & 1;
        mbar_wait(tempty_bar0 + 8 * acc, aph ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tb + (uint32_t)(acc * BNP);
        for (int kb = kb_begin; kb < kb_end; kb++, it++) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sA_hi = smem_base + s * STAGE_BYTES, sA_lo = sA_hi + TILE, sB_hi = sA_lo + TILE, sB_lo = sB_hi + TILE;
          const uint32_t a_hi = AK ? desc_lo_kmajor(sA_hi) : desc_lo_mnmajor(sA_hi), a_lo = AK ? desc_lo_kmajor(sA_lo) : desc_lo_mnmajor(sA_lo);
          const uint32_t b_hi = BKM ? desc_lo_kmajor(sB_hi) : desc_lo_mnmajor(sB_hi), b_lo = BKM ? desc_lo_kmajor(sB_lo) : desc_lo_mnmajor(sB_lo);
          if (elect_one()) {
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < BK / 16; k++) {
                umma2_bf16_lo(tmem_d, a_lo + A_STEP * k, b_hi + B_STEP * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                umma2_bf16_lo(tmem_d, a_hi + A_STEP * k, b_lo + B_STEP * k, idesc, 1u);
                umma2_bf16_lo(tmem_d, a_hi + A_STEP * k, b_hi + B_STEP * k, idesc, 1u);
              }
            }
            umma2_commit_mcast(empty_bar0 + 8 * s, (uint16_t)3);  // frees this stage in BOTH CTAs
          }
          __syncwarp();
        }
        if (elect_one()) umma2_commit_mcast(tfull_bar0 + 8 * acc, (uint16_t)3);  // accumulators complete in BOTH CTAs
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes x 256 columns) =====================
    const int quad = warp & 3;
    float* tb = epi_buf + quad * (EPI_WARP_BYTES / 4);
    const uint32_t stage_buf = smem_u32(tb);
    if (p.tma_epi && lane == 0) prefetch_tensormap(&tmC);
    uint32_t chunk_no = 0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    int local = 0;
    SegIter si(cidx, ncl, num_kb_total, p.sk_tile0, p.sk_tiles, p.dp_end);
    Seg sg;
    for (; si.next(sg); local++) {
      const int rem = sg.tile;
      const int m0 = ((rem % p.tiles_mp) * 2 + rank) * BM, n0 = (rem / p.tiles_mp) * BNP;
      const int acc = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      const bool split = sg.kb0 != 0 || sg.kb1 != num_kb_total;  // partial sum of a tile shared with other CTA pairs
      const bool use_bias = p.bias && sg.kb0 == 0;
      float bias_next = (use_bias && n0 + lane < p.N) ? p.bias[n0 + lane] : 0.f;
      mbar_wait(tfull_bar0 + 8 * acc, aph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BNP; c += 32) {
        uint32_t v[32];
        LRCN_TMEM_LD_32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BNP + c), v);
        const float bias_lane = bias_next;
        if (p.tma_epi && use_bias && c + 32 < BNP && n0 + c + 32 + lane < p.N) bias_next = p.bias[n0 + c + 32 + lane];
        tmem_ld_wait();
        if (c + 32 >= BNP) {  // accumulator stage drained: tell the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar0 + 8 * acc);
            else mbar_arrive_cluster(mapa_u32(tempty_bar0 + 8 * acc, 0));
          }
        }
        const int nb = n0 + c;
        if (nb >= p.N || (p.dbg & 1)) continue;
        if (p.tma_epi) {
          if (m0 + quad * 32 < p.M)
            epilogue_chunk_tma(&tmC, stage_buf + (chunk_no & 1u) * 4096u, v, bias_lane, use_bias, nb, m0 + quad * 32, split || p.beta, lane);
          chunk_no++;
          continue;
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
          *reinterpret_cast<float4*>(tb + lane * EPI_LD + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        const int mrow0 = m0 + quad * 32;
        if (vec_ok && !split && nb + 32 <= p.N) {
          const int cq = lane & 7, rsub = lane >> 3;
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
          const int n = nb + lane;
          const bool nok = n < p.N;
          const float bv = (use_bias && nok) ? p.bias[n] : 0.f;
#pragma unroll 4
          for (int rr = 0; rr < 32; rr++) {
            const int m = mrow0 + rr;
            if (m < p.M && nok) {
              float x = tb[rr * EPI_LD + lane] + bv;
              float* cp = p.C + (size_t)m * p.ldc + n;
              if (split) {
                atomicAdd(cp, x);
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
    if (p.tma_epi && lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    g2::tmem_dealloc2(tmem_base);
  }
}

static int g2_num_sms = 148;
static bool g_gemm2_ready = false;
int g_gemm_dbg = 0;
bool init_gemm2_sm100() {
  using namespace g2;
  cudaError_t e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_bf16x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { set_sm100_error((std::string("cudaFuncSetAttribute(gemm2): ") + cudaGetErrorString(e)).c_str()); return false; }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g2_num_sms, cudaDevAttrMultiProcessorCount, dev);
  g_gemm2_ready = true;
  return true;
}

// same contract as gemm_bf16x3 (kernels.cuh); the caller decides when the pair kernel pays off
static bool gemm2_launch(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N, int K, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
                         int lda, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, int n_split, const __nv_bfloat16* B2_hi,
                         const __nv_bfloat16* B2_lo, int ldb2, float* C, int ldc, bool beta, const float* bias, __nv_bfloat16* C_hi,
                         __nv_bfloat16* C_lo, bool c_zeroed) {
  using namespace g2;
  if (M <= 0 || N <= 0 || K <= 0) return true;
  const int tm = (M + BM - 1) / BM, tmp = (tm + 1) / 2, tn = (N + BNP - 1) / BNP;
  const int num_kb = (K + BK - 1) / BK;
  const int max_cl = g2_num_sms / 2;
  // Fewer tiles than CTA pairs: stream-K -- every pair takes an equal share of the (tile, k-block) iterations and partial
  // tiles are summed in L2 by TMA reduce-adds (fp32 atomics without the TMA epilogue); needs fp32-only output.
  // Between one and three waves of tiles: hybrid -- the whole waves run as whole tiles, the last partial wave as stream-K.
  const int tiles = tmp * tn;
  int sk_tile0 = 0, sk_tiles = 0, dp_end = tiles;
  int ncl = tiles < max_cl ? tiles : max_cl;
  bool zero_all = false;
  int zero_col0 = -1;  // hybrid: zero the output columns from here on
  static const bool no_hybrid = getenv("LRCN_GEMM_HYBRID") == nullptr;  // measured: no gain on the decoder's shapes (zero-fill + second epilogue), so opt-in
  if (!C_hi && tiles * 10 <= max_cl * 7 && num_kb >= 8) {  // (>= 70 % of the pairs busy with whole tiles: not worth the zero-fill + second epilogue)
    long long want = (long long)tiles * num_kb / 4;  // at least ~4 k-blocks per pair
    if (want > max_cl) want = max_cl;
    if (want > tiles) { ncl = (int)want; sk_tiles = tiles; dp_end = 0; zero_all = true; }
  } else if (!C_hi && !no_hybrid && tiles > max_cl && tiles < 3 * max_cl && (tiles % max_cl) != 0) {
    const int rem = tiles % max_cl;
    if ((long long)rem * num_kb >= 3ll * max_cl && rem * 10 <= max_cl * 7) {
      sk_tile0 = tiles - rem; sk_tiles = rem; dp_end = sk_tile0;
      zero_col0 = (sk_tile0 / tmp) * BNP;
    }
  }
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  bool ok = true;
  if (a_kmajor) ok = ok && get_tensor_map_bf16(&ta_hi, A_hi, K, M, lda, BM) && get_tensor_map_bf16(&ta_lo, A_lo, K, M, lda, BM);
  else          ok = ok && get_tensor_map_bf16(&ta_hi, A_hi, M, K, lda, BK) && get_tensor_map_bf16(&ta_lo, A_lo, M, K, lda, BK);
  const int N1 = n_split > 0 ? n_split : N, N2 = N - N1;
  if (b_kmajor) ok = ok && get_tensor_map_bf16(&tb_hi, B_hi, K, N1, ldb, BNH) && get_tensor_map_bf16(&tb_lo, B_lo, K, N1, ldb, BNH);
  else          ok = ok && get_tensor_map_bf16(&tb_hi, B_hi, N1, K, ldb, BK) && get_tensor_map_bf16(&tb_lo, B_lo, N1, K, ldb, BK);
  CUtensorMap tb2_hi = tb_hi, tb2_lo = tb_lo;
  if (ok && n_split > 0) {
    if (b_kmajor) ok = get_tensor_map_bf16(&tb2_hi, B2_hi, K, N2, ldb2, BNH) && get_tensor_map_bf16(&tb2_lo, B2_lo, K, N2, ldb2, BNH);
    else          ok = get_tensor_map_bf16(&tb2_hi, B2_hi, N2, K, ldb2, BK) && get_tensor_map_bf16(&tb2_lo, B2_lo, N2, K, ldb2, BK);
  }
  if (!ok) return false;
  const bool tma_epi = gemm_tma_epilogue_ok(C, ldc, beta, C_hi);
  CUtensorMap tc = ta_hi;
  if (tma_epi && !get_tensor_map_f32_out(&tc, C, N, M, ldc)) return false;
  if (zero_all && !beta && !c_zeroed) {
    if (ldc == N) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), s);
    else cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
  } else if (zero_col0 >= 0 && !beta && !c_zeroed) {
    cudaMemset2DAsync(C + zero_col0, (size_t)ldc * sizeof(float), 0, (size_t)(N - zero_col0) * sizeof(float), M, s);
  }
  Gemm2Params p;
  p.M = M; p.N = N; p.K = K; p.tiles_mp = tmp; p.tiles_n = tn; p.sk_tile0 = sk_tile0; p.sk_tiles = sk_tiles; p.dp_end = dp_end; p.n_split = n_split;
  p.C = C; p.ldc = ldc; p.bias = bias; p.beta = beta ? 1 : 0; p.C_hi = C_hi; p.C_lo = C_lo; p.dbg = g_gemm_dbg; p.tma_epi = tma_epi ? 1 : 0;
  const int grid = ncl * 2;
  if (a_kmajor && b_kmajor) launch_pdl(gemm2_bf16x3_kernel<true, true>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, s, ta_hi, ta_lo, tb_hi, tb_lo, tb2_hi, tb2_lo, tc, p);
  else if (a_kmajor && !b_kmajor) launch_pdl(gemm2_bf16x3_kernel<true, false>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, s, ta_hi, ta_lo, tb_hi, tb_lo, tb2_hi, tb2_lo, tc, p);
  else if (!a_kmajor && b_kmajor) launch_pdl(gemm2_bf16x3_kernel<false, true>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, s, ta_hi, ta_lo, tb_hi, tb_lo, tb2_hi, tb2_lo, tc, p);
  else launch_pdl(gemm2_bf16x3_kernel<false, false>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, s, ta_hi, ta_lo, tb_hi, tb_lo, tb2_hi, tb2_lo, tc, p);
  if (g_counter) g_counter->n++;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { set_sm100_error((std::string("gemm2_bf16x3 launch: ") + cudaGetErrorString(e)).c_str()); return false; }
  return true;
}

bool gemm2_bf16x3(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N, int K, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
                  int lda, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo, int ldb, float* C, int ldc, bool beta, const float* bias,
                  __nv_bfloat16* C_hi, __nv_bfloat16* C_lo, bool c_zeroed) {
  return gemm2_launch(s, a_kmajor, b_kmajor, M, N, K, A_hi, A_lo, lda, B_hi, B_lo, ldb, 0, nullptr, nullptr, 0, C, ldc, beta, bias, C_hi, C_lo,
                      c_zeroed);
}

// C[M][N1+N2] = A * [B1 | B2]: two B operands sharing A (the x-part and h-part of an LSTM weight gradient) in ONE launch.
// Applies when the split falls on a tile boundary; returns false WITHOUT launching otherwise (the caller issues two GEMMs).
bool gemm2_bf16x3_dualB(cudaStream_t s, bool a_kmajor, bool b_kmajor, int M, int N1, int N2, int K, const __nv_bfloat16* A_hi,
                        const __nv_bfloat16* A_lo, int lda, const __nv_bfloat16* B1_hi, const __nv_bfloat16* B1_lo, int ldb1,
                        const __nv_bfloat16* B2_hi, const __nv_bfloat16* B2_lo, int ldb2, float* C, int ldc, bool c_zeroed, bool* launched) {
  *launched = false;
  if (!g_gemm2_ready || N1 <= 0 || N2 <= 0 || (N1 % g2::BNP) != 0 || M < 256) return true;
  *launched = true;
  return gemm2_launch(s, a_kmajor, b_kmajor, M, N1 + N2, K, A_hi, A_lo, lda, B1_hi, B1_lo, ldb1, N1, B2_hi, B2_lo, ldb2, C, ldc, false, nullptr,
                      nullptr, nullptr, c_zeroed);
}

bool gemm2_bind_abort(unsigned int* host_flag) { return dev_abort_bind(host_flag) == cudaSuccess; }

}  // namespace lrcn
