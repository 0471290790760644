// dp_p2p.cu -- data-parallel gradient exchange over NVLink peer memory, without NCCL kernels.
//
// One process per GPU; every rank maps its peers' gradient arenas (CUDA IPC).  Rank r OWNS a contiguous 1/N shard of the
// arena: it loads that shard from every rank over NVLink (fixed rank order, so the sum is deterministic and identical
// everywhere), and stores the sum back into every rank's arena in place -- reduce-scatter and all-gather fused in one
// kernel with owner-computes semantics (nobody else touches an owner's elements, so no staging buffers are needed).
// Cross-GPU ordering is two flag barriers in peer memory (system-scope release / acquire):
//   B1  every rank's backward pass is complete  ->  exchange kernel  ->  B2  every rank's stores have landed.
// Why not NCCL here: its kernels compete for SMs with the persistent tcgen05 GEMM / LSTM kernels (which need their whole
// grid co-resident), so a "bucketed, overlapped" allreduce ends up serialising; measured in profiles/r01_progress.md.
#include "kernels.cuh"

#include <stdio.h>

namespace lrcn {

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {  // peer memory is cached in L1 only and L1 is not coherent: bypass it
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// one warp: lane q signals rank q and waits for rank q's signal.  flag_base selects one of the independent flag sets of the
// control block (barriers issued concurrently from two streams must not share flags or epoch counters)
__global__ void xgpu_barrier_kernel(P2PPeers peers, unsigned int* epoch_ctr, int flag_base) {
  __shared__ unsigned int epoch;
  if (threadIdx.x == 0) { epoch = *epoch_ctr + 1u; *epoch_ctr = epoch; }
  __syncwarp();
  const unsigned int e = epoch;
  const int q = threadIdx.x;
  __threadfence_system();
  if (q < peers.nranks) {
    st_release_sys(&peers.ctl[q]->flags[flag_base + peers.rank], e);
    const unsigned int* mine = &peers.ctl[peers.rank]->flags[flag_base + q];
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (dev_aborted()) break;
      if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died or never reached the barrier
        printf("lrcn dp_p2p: barrier timeout, rank %d waiting for rank %d (epoch %u)\n", peers.rank, q, e);
        dev_abort_set();  // flag + drain instead of a trap (kernels.cuh): the ABI reports a sticky error
        break;
      }
    }
  }
}

// in-place allreduce(sum) of the gradient arena by owner-computes shards + fp64 loss total
__global__ void __launch_bounds__(256) allreduce_p2p_kernel(P2PPeers peers, size_t begin4, size_t end4, double* loss_total) {
  const int N = peers.nranks;
  for (size_t i = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int p = 0; p < N; p++) {
      const float4 x = ld_peer_f4(reinterpret_cast<const float4*>(peers.g[p]) + i);
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    for (int p = 0; p < N; p++) reinterpret_cast<float4*>(peers.g[p])[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_total) {
    double t = 0.0;
    for (int p = 0; p < N; p++) {
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(&peers.ctl[p]->loss_partial) : "memory");
      t += v;
    }
    *loss_total = t;
  }
  __threadfence_system();  // my stores into peer memory are performed before this kernel ends (B2 follows)
}

// reduce-scatter + Adam + all-gather in one kernel: the owner of a shard sums its gradients over all ranks, updates ITS slice of
// m, v and w exactly like adam_kernel (same operation order: replicas stay bit-identical to the replicated update) and stores
// the new weights into every rank's arena.  Adam costs 1/N per GPU and the NVLink volume is the same as the allreduce's.
__global__ void __launch_bounds__(256) adam_p2p_kernel(P2PPeers peers, size_t begin4, size_t end4, float4* __restrict__ m, float4* __restrict__ v,
                                                       const StepScalars* __restrict__ sc, double* loss_total) {
  const int N = peers.nranks;
  const float b1 = sc->beta1, b2 = sc->beta2, lr = sc->lr, eps = sc->eps, d1 = sc->adam_d1, d2 = sc->adam_d2;
  const float ob1 = sc->one_m_beta1, ob2 = sc->one_m_beta2;
  const float4* wl = reinterpret_cast<const float4*>(peers.w[peers.rank]);
  for (size_t i = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += (size_t)gridDim.x * blockDim.x) {
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int p = 0; p < N; p++) {
      const float4 x = ld_peer_f4(reinterpret_cast<const float4*>(peers.g[p]) + i);
      G.x += x.x; G.y += x.y; G.z += x.z; G.w += x.w;
    }
    float4 W = wl[i], Mv = m[i], Vv = v[i];
    float* wp = reinterpret_cast<float*>(&W);
    const float* gp = reinterpret_cast<const float*>(&G);
    float* mp = reinterpret_cast<float*>(&Mv);
    float* vp = reinterpret_cast<float*>(&Vv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float mm = __fadd_rn(__fmul_rn(b1, mp[k]), __fmul_rn(ob1, gp[k]));
      const float vv = __fadd_rn(__fmul_rn(b2, vp[k]), __fmul_rn(ob2, __fmul_rn(gp[k], gp[k])));
      const float upd = __fdiv_rn(__fdiv_rn(mm, d1), __fadd_rn(__fsqrt_rn(__fdiv_rn(vv, d2)), eps));
      wp[k] = __fsub_rn(wp[k], __fmul_rn(lr, upd));
      mp[k] = mm; vp[k] = vv;
    }
    m[i] = Mv; v[i] = Vv;
    for (int p = 0; p < N; p++) reinterpret_cast<float4*>(peers.w[p])[i] = W;
    reinterpret_cast<float4*>(peers.g[peers.rank])[i] = G;  // the summed gradient of the owned shard stays readable (lrcn_get_grad)
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_total) {
    double t = 0.0;
    for (int p = 0; p < N; p++) {
      double x;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(&peers.ctl[p]->loss_partial) : "memory");
      t += x;
    }
    *loss_total = t;
  }
  __threadfence_system();
}

bool dp_bind_abort(unsigned int* host_flag) { return dev_abort_bind(host_flag) == cudaSuccess; }

void dp_p2p_shard(size_t n_floats, int nranks, int r, size_t* begin, size_t* end) {
  const size_t n4 = n_floats / 4, per = (n4 + nranks - 1) / nranks;
  const size_t b = per * r < n4 ? per * r : n4;
  const size_t e = b + per < n4 ? b + per : n4;
  *begin = 4 * b; *end = 4 * e;
}

void dp_p2p_adam(cudaStream_t s, const P2PPeers& peers, size_t n_floats, unsigned int* epoch_ctr, double* loss_total, float* m, float* v,
                 const StepScalars* sc) {
  size_t b, e;
  dp_p2p_shard(n_floats, peers.nranks, peers.rank, &b, &e);
  xgpu_barrier_kernel<<<1, 32, 0, s>>>(peers, epoch_ctr, 0);
  adam_p2p_kernel<<<148 * 4, 256, 0, s>>>(peers, b / 4, e / 4, reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), sc, loss_total);
  xgpu_barrier_kernel<<<1, 32, 0, s>>>(peers, epoch_ctr, 0);
  if (g_counter) g_counter->n += 3;
}

void dp_p2p_allreduce(cudaStream_t s, const P2PPeers& peers, size_t n_floats, unsigned int* epoch_ctr, double* loss_total) {
  size_t b, e;  // arena sizes are multiples of 64 floats
  dp_p2p_shard(n_floats, peers.nranks, peers.rank, &b, &e);
  xgpu_barrier_kernel<<<1, 32, 0, s>>>(peers, epoch_ctr, 0);
  allreduce_p2p_kernel<<<148 * 4, 256, 0, s>>>(peers, b / 4, e / 4, loss_total);
  xgpu_barrier_kernel<<<1, 32, 0, s>>>(peers, epoch_ctr, 0);
  if (g_counter) g_counter->n += 3;
}

// ---------------------------------------------------------------------------------------------------------------------
// Copy-engine exchange (default for training).  The NVLink traffic of the gradient reduce-scatter and of the weight
// all-gather is done by cudaMemcpyAsync nodes (DMA engines, no SMs): every rank PUSHES the slices of a gradient bucket into
// their owners' staging buffers, a flag barrier says "my pushes have landed", the owner sums the N contributions and runs
// Adam on its slice with the small local kernel below, and pushes the new weights back.  Because no SM is involved in the
// transfers, the buckets that are ready early (Wout/bout after the vocab backward; W2, b2, Wf, Wcnn after layer-2 BPTT) are
// exchanged on a side stream UNDER the persistent LSTM / GEMM kernels of the rest of the backward pass, which the SM-driven
// exchange (and NCCL's kernels) could not do without stealing SMs from kernels that need their whole grid resident.
// stage: [nranks][stride4] float4 staging rows on the owner (row q = contribution of rank q; the owner's own row is unused,
// its contribution is read from g directly).  Same operation order as adam_kernel: replicas stay bit-identical.
__global__ void __launch_bounds__(256) adam_staged_kernel(float4* __restrict__ w, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                                          const float4* __restrict__ stage, size_t stride4, size_t stage_off4, size_t b4, size_t e4, int nranks,
                                                          int rank, const StepScalars* __restrict__ sc, P2PPeers peers, double* loss_total) {
  const float b1 = sc->beta1, b2 = sc->beta2, lr = sc->lr, eps = sc->eps, d1 = sc->adam_d1, d2 = sc->adam_d2;
  const float ob1 = sc->one_m_beta1, ob2 = sc->one_m_beta2;
  for (size_t i = b4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < e4; i += (size_t)gridDim.x * blockDim.x) {
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t so = stage_off4 + (i - b4);
    for (int q = 0; q < nranks; q++) {  // fixed rank order: the sum is deterministic and identical to the SM-driven exchange
      const float4 x = q == rank ? g[i] : __ldcs(stage + (size_t)q * stride4 + so);
      G.x += x.x; G.y += x.y; G.z += x.z; G.w += x.w;
    }
    float4 W = w[i], Mv = m[i], Vv = v[i];
    float* wp = reinterpret_cast<float*>(&W);
    const float* gp = reinterpret_cast<const float*>(&G);
    float* mp = reinterpret_cast<float*>(&Mv);
    float* vp = reinterpret_cast<float*>(&Vv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float mm = __fadd_rn(__fmul_rn(b1, mp[k]), __fmul_rn(ob1, gp[k]));
      const float vv = __fadd_rn(__fmul_rn(b2, vp[k]), __fmul_rn(ob2, __fmul_rn(gp[k], gp[k])));
      const float upd = __fdiv_rn(__fdiv_rn(mm, d1), __fadd_rn(__fsqrt_rn(__fdiv_rn(vv, d2)), eps));
      wp[k] = __fsub_rn(wp[k], __fmul_rn(lr, upd));
      mp[k] = mm; vp[k] = vv;
    }
    w[i] = W; m[i] = Mv; v[i] = Vv;
    g[i] = G;  // the summed gradient of the owned slice stays readable (lrcn_get_grad gathers the slices)
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss_total) {
    double t = 0.0;
    for (int p = 0; p < nranks; p++) {
      double x;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(&peers.ctl[p]->loss_partial) : "memory");
      t += x;
    }
    *loss_total = t;
  }
}
void dp_adam_staged(cudaStream_t s, float* w, float* g, float* m, float* v, const float* stage, size_t stride, size_t stage_off, size_t b, size_t e, const P2PPeers& peers,
                    const StepScalars* sc, double* loss_total) {
  const size_t n4 = (e - b) / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 2) grid = 148 * 2;
  if (grid < 1) grid = 1;
  adam_staged_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(w), reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
                                          reinterpret_cast<const float4*>(stage), stride / 4, stage_off / 4, b / 4, e / 4, peers.nranks, peers.rank, sc, peers, loss_total);
  if (g_counter) g_counter->n++;
}
// SM-driven exchange of ONE arena range [b, e) owned by this rank (reduce-scatter + Adam + all-gather in one kernel, no
// barriers: the caller brackets it with dp_xgpu_barrier).  grid_ctas bounds the SMs it may take.
void dp_p2p_adam_range(cudaStream_t s, const P2PPeers& peers, size_t b, size_t e, float* m, float* v, const StepScalars* sc, double* loss_total, int grid_ctas) {
  adam_p2p_kernel<<<grid_ctas, 256, 0, s>>>(peers, b / 4, e / 4, reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), sc, loss_total);
  if (g_counter) g_counter->n++;
}
void dp_xgpu_barrier(cudaStream_t s, const P2PPeers& peers, unsigned int* epoch_ctr, int flagset) {
  xgpu_barrier_kernel<<<1, 32, 0, s>>>(peers, epoch_ctr, 16 * flagset);
  if (g_counter) g_counter->n++;
}

}  // namespace lrcn
