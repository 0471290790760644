// lstm_sm100.cu -- one LSTM timestep as ONE tcgen05 kernel: recurrent gate GEMM (bf16 hi/lo split, fp32 TMEM
// accumulate) fused with the sigmoid/tanh cell update (forward) or with the cell adjoint (backward).
// Reference semantics: lstm() lrcn.jl:528-538 (gate order [forget|ingate|outgate|change]) and its adjoint
// (SURVEY.md §10.2).  The x-part of the gates (input projection + bias) is precomputed for all timesteps by
// one large GEMM (teacher forcing), so a step only has the h_{t-1} * W_h product left.
//
// forward  (per step t, layer with hidden size H, batch B):
//   tile = 128 batch rows x (4 gates x 16 hidden units); B operand = W_h rows permuted so that a CTA's 64 columns
//   are [f(16) i(16) o(16) g(16)] of the same 16 units -> a thread (= batch row, TMEM lane) owns all four gates
//   of its 16 units and finishes c_t, h_t in registers.  grid = (H/16, B/128).
// backward (per step t): dh_rec^T[j][m] = sum_n W_h[n][j] * dG_{t+1}[m][n]   ("swap-AB": M = 128 hidden units from
//   the MN-major weight view, N = 32 batch rows), epilogue thread = hidden unit j: dh = dh_in + dh_rec, cell adjoint,
//   writes dG_t (fp32 in place over the stored activations + bf16 hi/lo for the next step) and dc.  grid = (H/128, B/32).
//   All epilogue global accesses are coalesced over j.
#include "kernels.cuh"
#include "sm100_ptx.cuh"

#include <string>

namespace lrcn {
using namespace ptx;

constexpr int LBK = 64;
constexpr int L_A_TILE = 128 * LBK * 2;  // 16 KiB
constexpr int L_THREADS = 192;

// ---------------------------------------------------------------------------------------------- forward
constexpr int F_NH = 16, F_BN = 4 * F_NH;            // 64 gate columns per CTA
constexpr int F_B_TILE = F_BN * LBK * 2;             // 8 KiB
constexpr int F_STAGE = 2 * L_A_TILE + 2 * F_B_TILE; // 48 KiB
constexpr int F_STAGES = 4;
constexpr int F_SMEM = F_STAGES * F_STAGE + 1024 + 256;

struct LstmFwdParams {
  int B, H, num_kb, has_rec;
  float* gates;         // [B][4H] pre-activations (x-part + bias) in, activations out
  const float* c_prev;  // [B][H]
  float* c_out;         // [B][H]
  float* h_out;         // [B][H]
  __nv_bfloat16* h_hi;  // bf16 split of h_out (A operand of the next step and of the batched GEMMs), may be null
  __nv_bfloat16* h_lo;
};

__device__ __forceinline__ float sigm_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(L_THREADS, 1)
lstm_fwd_step_kernel(const __grid_constant__ CUtensorMap tmH_hi, const __grid_constant__ CUtensorMap tmH_lo,
                     const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const LstmFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + F_STAGES * F_STAGE);
  const uint32_t full_bar0 = smem_u32(bars), empty_bar0 = smem_u32(bars + F_STAGES), tfull_bar = smem_u32(bars + 2 * F_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * F_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, m0 = blockIdx.y * 128;
  const int num_kb = p.has_rec ? p.num_kb : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < F_STAGES; s++) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, 1); }
    mbar_init(tfull_bar, 1);
    mbar_init_fence();
  }
  if (warp == 0 && lane == 0 && num_kb > 0) {
    prefetch_tensormap(&tmH_hi); prefetch_tensormap(&tmH_lo); prefetch_tensormap(&tmW_hi); prefetch_tensormap(&tmW_lo);
  }
  if (warp == 1) tmem_alloc<F_BN>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; i++) {
        const int s = i % F_STAGES;
        mbar_wait(empty_bar0 + 8 * s, ((i / F_STAGES) & 1) ^ 1);
        const uint32_t full = full_bar0 + 8 * s;
        mbar_expect_tx(full, F_STAGE);
        const uint32_t st = smem_base + s * F_STAGE;
        tma_load_2d(st, &tmH_hi, full, i * LBK, m0);
        tma_load_2d(st + L_A_TILE, &tmH_lo, full, i * LBK, m0);
        tma_load_2d(st + 2 * L_A_TILE, &tmW_hi, full, i * LBK, jt * F_BN);
        tma_load_2d(st + 2 * L_A_TILE + F_B_TILE, &tmW_lo, full, i * LBK, jt * F_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && num_kb > 0) {
      const uint32_t idesc = idesc_bf16(128, F_BN, false, false);
      for (int i = 0; i < num_kb; i++) {
        const int s = i % F_STAGES;
        mbar_wait(full_bar0 + 8 * s, (i / F_STAGES) & 1);
        tc_fence_after();
        const uint32_t st = smem_base + s * F_STAGE;
#pragma unroll
        for (int k = 0; k < LBK / 16; k++) {
          const uint64_t a_hi = desc_kmajor(st, k), a_lo = desc_kmajor(st + L_A_TILE, k);
          const uint64_t b_hi = desc_kmajor(st + 2 * L_A_TILE, k), b_lo = desc_kmajor(st + 2 * L_A_TILE + F_B_TILE, k);
          umma_bf16(tmem_base, a_lo, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_bf16(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(empty_bar0 + 8 * s);
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + lane;
    uint32_t v[64];
    if (num_kb > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      LRCN_TMEM_LD_32(tmem_base + ((uint32_t)(quad * 32) << 16), v);
      { uint32_t* v2 = v + 32; LRCN_TMEM_LD_32(tmem_base + ((uint32_t)(quad * 32) << 16) + 32u, v2); }
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 64; j++) v[j] = 0u;
    }
    const int H = p.H, j0 = jt * F_NH;
    if (m < p.B) {
      float* grow = p.gates + (size_t)m * 4 * H;
      const size_t hidx = (size_t)m * H + j0;
#pragma unroll
      for (int q = 0; q < F_NH / 4; q++) {  // 4 units at a time (H % 4 == 0 -> a float4 never straddles H)
        const int j = j0 + 4 * q;
        if (j < H) {
          float4 gf = *reinterpret_cast<const float4*>(grow + j);
          float4 gi = *reinterpret_cast<const float4*>(grow + H + j);
          float4 go = *reinterpret_cast<const float4*>(grow + 2 * H + j);
          float4 gg = *reinterpret_cast<const float4*>(grow + 3 * H + j);
          const float4 cp = *reinterpret_cast<const float4*>(p.c_prev + hidx + 4 * q);
          float f[4] = {gf.x, gf.y, gf.z, gf.w}, in[4] = {gi.x, gi.y, gi.z, gi.w}, o[4] = {go.x, go.y, go.z, go.w}, ch[4] = {gg.x, gg.y, gg.z, gg.w};
          const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
          float cn[4], hn[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int u = 4 * q + e;
            f[e] = sigm_f(f[e] + __uint_as_float(v[u]));
            in[e] = sigm_f(in[e] + __uint_as_float(v[F_NH + u]));
            o[e] = sigm_f(o[e] + __uint_as_float(v[2 * F_NH + u]));
            ch[e] = tanhf(ch[e] + __uint_as_float(v[3 * F_NH + u]));
            cn[e] = cpv[e] * f[e] + in[e] * ch[e];
            hn[e] = o[e] * tanhf(cn[e]);
          }
          *reinterpret_cast<float4*>(grow + j) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(grow + H + j) = make_float4(in[0], in[1], in[2], in[3]);
          *reinterpret_cast<float4*>(grow + 2 * H + j) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(grow + 3 * H + j) = make_float4(ch[0], ch[1], ch[2], ch[3]);
          *reinterpret_cast<float4*>(p.c_out + hidx + 4 * q) = make_float4(cn[0], cn[1], cn[2], cn[3]);
          *reinterpret_cast<float4*>(p.h_out + hidx + 4 * q) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          if (p.h_hi) {
            __nv_bfloat16 hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; e++) split_bf16(hn[e], hh[e], ll[e]);
            *reinterpret_cast<uint2*>(p.h_hi + hidx + 4 * q) = *reinterpret_cast<uint2*>(hh);
            *reinterpret_cast<uint2*>(p.h_lo + hidx + 4 * q) = *reinterpret_cast<uint2*>(ll);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<F_BN>(tmem_base);
  }
}

// W_h (rows = gate columns n = g*H + j of the reference's (X+H) x 4H weight, K-major with pitch ldw, columns
// [x_off, x_off+H)) -> permuted bf16 hi/lo [ceil(H/16)*64][Hp]: row (jt*4 + g)*16 + u  <-  n = g*H + jt*16 + u
__global__ void permute_split_kernel(const float* __restrict__ W, int ldw, int x_off, int H, int Hp, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo) {
  const int row = blockIdx.x;  // permuted row
  const int jt = row / F_BN, r = row % F_BN, g = r / F_NH, u = r % F_NH;
  const int j = jt * F_NH + u;
  for (int k = threadIdx.x; k < Hp; k += blockDim.x) {
    float x = 0.f;
    if (j < H && k < H) x = W[(size_t)(g * H + j) * ldw + x_off + k];
    __nv_bfloat16 h, l;
    split_bf16(x, h, l);
    hi[(size_t)row * Hp + k] = h;
    lo[(size_t)row * Hp + k] = l;
  }
}
void lstm_permute_weights(cudaStream_t s, const float* W, int ldw, int x_off, int H, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const int rows = (H + F_NH - 1) / F_NH * F_BN;
  const int Hp = (H + 7) / 8 * 8;
  permute_split_kernel<<<rows, 128, 0, s>>>(W, ldw, x_off, H, Hp, hi, lo);
  if (g_counter) g_counter->n++;
}
size_t lstm_permuted_elems(int H) { return (size_t)((H + F_NH - 1) / F_NH * F_BN) * ((H + 7) / 8 * 8); }

bool lstm_fwd_step(cudaStream_t s, int B, int H, bool has_rec, const __nv_bfloat16* hprev_hi, const __nv_bfloat16* hprev_lo,
                   const __nv_bfloat16* wperm_hi, const __nv_bfloat16* wperm_lo, float* gates, const float* c_prev, float* c_out, float* h_out,
                   __nv_bfloat16* h_hi, __nv_bfloat16* h_lo) {
  const int Hp = (H + 7) / 8 * 8;
  const int rows = (H + F_NH - 1) / F_NH * F_BN;
  CUtensorMap th_hi, th_lo, tw_hi, tw_lo;
  if (has_rec) {
    if (!get_tensor_map_bf16(&th_hi, hprev_hi, H, B, H, 128) || !get_tensor_map_bf16(&th_lo, hprev_lo, H, B, H, 128)) return false;
  } else {  // never dereferenced, but the kernel parameters must be valid maps
    if (!get_tensor_map_bf16(&th_hi, wperm_hi, H, rows, Hp, 128) || !get_tensor_map_bf16(&th_lo, wperm_lo, H, rows, Hp, 128)) return false;
  }
  if (!get_tensor_map_bf16(&tw_hi, wperm_hi, H, rows, Hp, F_BN) || !get_tensor_map_bf16(&tw_lo, wperm_lo, H, rows, Hp, F_BN)) return false;
  LstmFwdParams p;
  p.B = B; p.H = H; p.num_kb = (H + LBK - 1) / LBK; p.has_rec = has_rec ? 1 : 0;
  p.gates = gates; p.c_prev = c_prev; p.c_out = c_out; p.h_out = h_out; p.h_hi = h_hi; p.h_lo = h_lo;
  dim3 grid((H + F_NH - 1) / F_NH, (B + 127) / 128);
  lstm_fwd_step_kernel<<<grid, L_THREADS, F_SMEM, s>>>(th_hi, th_lo, tw_hi, tw_lo, p);
  if (g_counter) g_counter->n++;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { set_sm100_error((std::string("lstm_fwd_step launch: ") + cudaGetErrorString(e)).c_str()); return false; }
  return true;
}

// ---------------------------------------------------------------------------------------------- backward
constexpr int R_BN = 32;                              // batch rows per CTA (UMMA N)
constexpr int R_B_TILE = R_BN * LBK * 2;              // 4 KiB
constexpr int R_STAGE = 2 * L_A_TILE + 2 * R_B_TILE;  // 40 KiB
constexpr int R_STAGES = 5;
constexpr int R_SMEM = R_STAGES * R_STAGE + 1024 + 256;

struct LstmBwdParams {
  int B, H, num_kb, has_rec;
  float* gates;          // [B][4H]: activations (f,i,o,g) in, dG out
  __nv_bfloat16* g_hi;   // bf16 split of dG (B operand of the next step, A/B operand of the batched gradient GEMMs)
  __nv_bfloat16* g_lo;
  const float* c_prev;   // c_{t-1}
  const float* c_cur;    // c_t
  const float* dh_in;    // dL/dh_t from the layer above
  float* dc;             // carry: in dL/dc_t (from step t+1), out dL/dc_{t-1}
};

__global__ void __launch_bounds__(L_THREADS, 1)
lstm_bwd_step_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                     const __grid_constant__ CUtensorMap tmG_hi, const __grid_constant__ CUtensorMap tmG_lo, const LstmBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + R_STAGES * R_STAGE);
  const uint32_t full_bar0 = smem_u32(bars), empty_bar0 = smem_u32(bars + R_STAGES), tfull_bar = smem_u32(bars + 2 * R_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * 128, m0 = blockIdx.y * R_BN;
  const int num_kb = p.has_rec ? p.num_kb : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R_STAGES; s++) { mbar_init(full_bar0 + 8 * s, 1); mbar_init(empty_bar0 + 8 * s, 1); }
    mbar_init(tfull_bar, 1);
    mbar_init_fence();
  }
  if (warp == 0 && lane == 0 && num_kb > 0) {
    prefetch_tensormap(&tmW_hi); prefetch_tensormap(&tmW_lo); prefetch_tensormap(&tmG_hi); prefetch_tensormap(&tmG_lo);
  }
  if (warp == 1) tmem_alloc<R_BN>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; i++) {
        const int s = i % R_STAGES;
        mbar_wait(empty_bar0 + 8 * s, ((i / R_STAGES) & 1) ^ 1);
        const uint32_t full = full_bar0 + 8 * s;
        mbar_expect_tx(full, R_STAGE);
        const uint32_t st = smem_base + s * R_STAGE;
        const int k0 = i * LBK;  // k runs over the 4H gate columns
#pragma unroll
        for (int b = 0; b < 2; b++) {  // A = W_h viewed MN-major: boxes [64 k][64 j]
          tma_load_2d(st + b * 8192, &tmW_hi, full, j0 + 64 * b, k0);
          tma_load_2d(st + L_A_TILE + b * 8192, &tmW_lo, full, j0 + 64 * b, k0);
        }
        tma_load_2d(st + 2 * L_A_TILE, &tmG_hi, full, k0, m0);  // B = dG_{t+1} [32 m][64 k]
        tma_load_2d(st + 2 * L_A_TILE + R_B_TILE, &tmG_lo, full, k0, m0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && num_kb > 0) {
      const uint32_t idesc = idesc_bf16(128, R_BN, true, false);
      for (int i = 0; i < num_kb; i++) {
        const int s = i % R_STAGES;
        mbar_wait(full_bar0 + 8 * s, (i / R_STAGES) & 1);
        tc_fence_after();
        const uint32_t st = smem_base + s * R_STAGE;
#pragma unroll
        for (int k = 0; k < LBK / 16; k++) {
          const uint64_t a_hi = desc_mnmajor(st, k), a_lo = desc_mnmajor(st + L_A_TILE, k);
          const uint64_t b_hi = desc_kmajor(st + 2 * L_A_TILE, k), b_lo = desc_kmajor(st + 2 * L_A_TILE + R_B_TILE, k);
          umma_bf16(tmem_base, a_lo, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_bf16(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(empty_bar0 + 8 * s);
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quad = warp & 3;
    const int j = j0 + quad * 32 + lane;  // this thread's hidden unit (TMEM lane)
    uint32_t v[32];
    if (num_kb > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      LRCN_TMEM_LD_32(tmem_base + ((uint32_t)(quad * 32) << 16), v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int c = 0; c < 32; c++) v[c] = 0u;
    }
    const int H = p.H;
    if (j < H) {
#pragma unroll
      for (int c = 0; c < R_BN; c++) {
        const int m = m0 + c;
        if (m < p.B) {
          const size_t idx = (size_t)m * H + j;
          float* g = p.gates + (size_t)m * 4 * H + j;
          const float f = g[0], in = g[H], o = g[2 * H], ch = g[3 * H];
          const float dh = p.dh_in[idx] + __uint_as_float(v[c]);
          const float tc = tanhf(p.c_cur[idx]);
          const float dcv = (p.has_rec ? p.dc[idx] : 0.f) + dh * o * (1.f - tc * tc);
          const float dO = dh * tc, dF = dcv * p.c_prev[idx], dI = dcv * ch, dG = dcv * in;
          p.dc[idx] = dcv * f;
          const float r0 = dF * f * (1.f - f), r1 = dI * in * (1.f - in), r2 = dO * o * (1.f - o), r3 = dG * (1.f - ch * ch);
          g[0] = r0; g[H] = r1; g[2 * H] = r2; g[3 * H] = r3;
          if (p.g_hi) {
            __nv_bfloat16* gh = p.g_hi + (size_t)m * 4 * H + j;
            __nv_bfloat16* gl = p.g_lo + (size_t)m * 4 * H + j;
            __nv_bfloat16 hh, ll;
            split_bf16(r0, hh, ll); gh[0] = hh; gl[0] = ll;
            split_bf16(r1, hh, ll); gh[H] = hh; gl[H] = ll;
            split_bf16(r2, hh, ll); gh[2 * H] = hh; gl[2 * H] = ll;
            split_bf16(r3, hh, ll); gh[3 * H] = hh; gl[3 * H] = ll;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<R_BN>(tmem_base);
  }
}

// w_hi/w_lo: bf16 shadows of the layer's full weight [4H][ldw]; the recurrent block is columns [x_off, x_off+H)
bool lstm_bwd_step(cudaStream_t s, int B, int H, bool has_rec, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int ldw, int x_off,
                   const __nv_bfloat16* gnext_hi, const __nv_bfloat16* gnext_lo, float* gates, __nv_bfloat16* g_hi, __nv_bfloat16* g_lo,
                   const float* c_prev, const float* c_cur, const float* dh_in, float* dc) {
  CUtensorMap tw_hi, tw_lo, tg_hi, tg_lo;
  if (!get_tensor_map_bf16(&tw_hi, w_hi + x_off, H, 4 * (uint64_t)H, ldw, 64) || !get_tensor_map_bf16(&tw_lo, w_lo + x_off, H, 4 * (uint64_t)H, ldw, 64))
    return false;
  const __nv_bfloat16* gh = has_rec ? gnext_hi : g_hi;  // a valid map is needed even when unused
  const __nv_bfloat16* gl = has_rec ? gnext_lo : g_lo;
  if (!get_tensor_map_bf16(&tg_hi, gh, 4 * (uint64_t)H, B, 4 * (uint64_t)H, R_BN) || !get_tensor_map_bf16(&tg_lo, gl, 4 * (uint64_t)H, B, 4 * (uint64_t)H, R_BN))
    return false;
  LstmBwdParams p;
  p.B = B; p.H = H; p.num_kb = (4 * H + LBK - 1) / LBK; p.has_rec = has_rec ? 1 : 0;
  p.gates = gates; p.g_hi = g_hi; p.g_lo = g_lo; p.c_prev = c_prev; p.c_cur = c_cur; p.dh_in = dh_in; p.dc = dc;
  dim3 grid((H + 127) / 128, (B + R_BN - 1) / R_BN);
  lstm_bwd_step_kernel<<<grid, L_THREADS, R_SMEM, s>>>(tw_hi, tw_lo, tg_hi, tg_lo, p);
  if (g_counter) g_counter->n++;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { set_sm100_error((std::string("lstm_bwd_step launch: ") + cudaGetErrorString(e)).c_str()); return false; }
  return true;
}

bool init_lstm_sm100() {
  cudaError_t e = cudaFuncSetAttribute(lstm_fwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM);
  if (e != cudaSuccess) { set_sm100_error((std::string("cudaFuncSetAttribute(lstm): ") + cudaGetErrorString(e)).c_str()); return false; }
  return true;
}

}  // namespace lrcn
