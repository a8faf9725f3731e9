// Data-parallel optimiser step over NVLink / NVSwitch peer memory (the one collective of the path).
//
// The reference's only multi-GPU strategy is DDP (fruit_pipeline.py:119-121): every rank back-propagates its own rays, the
// gradients are all-reduced (mean) and every rank runs the same Adam step (fruit_nerf_config.py:45-60) on its replica.  Done
// with NCCL that is: all-reduce (2(N-1)/N x 77.6 MB on the wire per GPU) followed by a dense Adam pass (7 x 4 B x 19.4 M
// parameters of HBM traffic on EVERY rank).  Here both are ONE kernel over peer-mapped (symmetric) buffers:
//
//   rank r owns the slice [r n/N, (r+1) n/N) of every flat parameter group.  For each float4 of its slice it
//     1. loads the gradient from all N ranks (peer LDG over NVLink, fixed rank order -> every replica gets the same bits),
//     2. applies Adam with ITS shard of the moments (exp_avg / exp_avg_sq are only maintained for the owned slice:
//        optimiser HBM traffic and state drop by N, ZeRO-1 style),
//     3. stores the updated parameters into all N replicas (peer STG).
//   = reduce-scatter + Adam + all-gather with the same wire bytes as the ring all-reduce, no staging buffers, no second pass;
//   inbound gradient reads and outbound parameter writes use the two directions of the links concurrently.
//
// With NVLS (multicast-mapped buffers) the N loads become one multimem.ld_reduce (the switch adds) and the N stores one
// multimem.st (the switch replicates): CNB_P2P_MULTIMEM.
//
// Cross-GPU ordering: k_p2p_barrier (one warp): every rank publishes a sequence number into its slot of every peer's flag block
// with a system-scope release and spins (acquire) until all peers published theirs.  barrier -> update kernel(s) -> barrier ->
// each rank clears its own gradient.  A spin that exceeds the timeout sets state[1] and gives up, so a dead peer cannot hang the GPU.
#include "cnb_common.cuh"

namespace {

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(32) k_p2p_barrier(cnb_p2p_comm c, uint64_t timeout_ns) {
  __shared__ uint32_t seq_s;
  const int ch = c.channel;
  if (threadIdx.x == 0) seq_s = ++c.state[2 * ch];  // stream-ordered: only this channel's barriers touch the counter
  __syncwarp();
  const uint32_t seq = seq_s;
  const int k = threadIdx.x;
  if (k < c.world) {
    __threadfence_system();
    st_release_sys(c.flags[k] + ch * CNB_MAX_PEERS + c.rank, seq);
    const uint32_t* mine = c.flags[c.rank] + ch * CNB_MAX_PEERS + k;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(c.state + 2 * ch + 1, 1u); break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
}

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 mm_ld_reduce4(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void mm_st4(float4* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct AdamArgs {
  float lr, b1, b2, eps, bc1, bc2_sqrt, inv_scale;
};

// MODE 0: peer loads / stores; MODE 1: NVLS multimem.  WORLD > 0: compile-time rank count; 0: run-time (<= CNB_MAX_PEERS).
// Every thread issues the WORLD peer loads of its element (x U elements) before the first add, then its moments / parameters, updates and scatters
// the new parameters.  Measured at N=2 (tests/ddp_p2p_check.py): U=1 at 45 registers (5 CTAs/SM) beats U=4 at 128 registers (0.14 vs 0.16 ms for the
// two flat groups): occupancy, not per-thread unrolling, is what hides the 2-4 us peer latency; the specialised rank counts all use U=1.
template <int MODE, int WORLD, int U>
__global__ void __launch_bounds__(256, (U * (WORLD > 0 ? WORLD : CNB_MAX_PEERS) <= 2) ? 5 : ((U * (WORLD > 0 ? WORLD : CNB_MAX_PEERS) <= 4) ? 4 : 2)) k_ddp_adam(cnb_p2p_comm c, cnb_p2p_group g, float4* __restrict__ m, float4* __restrict__ v, int64_t lo4, int64_t hi4,
                                                  AdamArgs a, int grads_zero, const float* __restrict__ dev_scalars) {
  if (dev_scalars != nullptr) {  // graph-replayed step: this step's scalars live in device memory (cnb_opt_group.scalars layout)
    a.lr = __ldg(dev_scalars); a.b1 = __ldg(dev_scalars + 1); a.b2 = __ldg(dev_scalars + 2); a.eps = __ldg(dev_scalars + 3);
    a.bc1 = __ldg(dev_scalars + 4); a.bc2_sqrt = __ldg(dev_scalars + 5); a.inv_scale = __ldg(dev_scalars + 6);
  }
  const float step_size = a.lr / a.bc1;
  const int world = WORLD > 0 ? WORLD : c.world;
  constexpr int NP = WORLD > 0 ? WORLD : CNB_MAX_PEERS;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4* __restrict__ p_own = reinterpret_cast<const float4*>(g.param[c.rank]);
  const uint32_t* __restrict__ live = g.live;
  for (int64_t base = lo4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; base < hi4; base += stride * U) {
    // unreachable hash-table rows (cnb_hashgrid_mark_reachable; the bitmap is the same on every rank): zero gradient and moments on all ranks
    // forever -> nothing to reduce, update or broadcast
    if (U == 1 && live != nullptr && !((__ldg(live + (base >> 5)) >> (base & 31)) & 1u)) continue;
    float4 gs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) gs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(grads_zero & 1)) {
      if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t i = base + u * stride;
          if (i < hi4) gs[u] = mm_ld_reduce4(reinterpret_cast<const float4*>(g.mc_grad) + i);
        }
      } else {
        float4 part[U][NP];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t i = base + u * stride;
#pragma unroll
          for (int k = 0; k < NP; ++k)
            if (k < world && i < hi4) part[u][k] = ld_stream4(reinterpret_cast<const float4*>(g.grad[(grads_zero & 2) ? c.rank : k]) + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t i = base + u * stride;
#pragma unroll
          for (int k = 0; k < NP; ++k)  // rank order: the same bits on every replica
            if (k < world && i < hi4) { gs[u].x += part[u][k].x; gs[u].y += part[u][k].y; gs[u].z += part[u][k].z; gs[u].w += part[u][k].w; }
        }
      }
    }
    float4 mi[U], vi[U], pi[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + u * stride;
      if (i < hi4) { mi[u] = m[i]; vi[u] = v[i]; pi[u] = p_own[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + u * stride;
      if (i >= hi4) continue;
      float* gp = &gs[u].x; float* mp = &mi[u].x; float* vp = &vi[u].x; float* pp = &pi[u].x;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float gk = gp[q] * a.inv_scale;
        mp[q] = mp[q] + (1.0f - a.b1) * (gk - mp[q]);
        vp[q] = a.b2 * vp[q] + (1.0f - a.b2) * gk * gk;
        pp[q] -= step_size * (mp[q] / (sqrtf(vp[q]) / a.bc2_sqrt + a.eps));
      }
      m[i] = mi[u]; v[i] = vi[u];
      if (MODE == 1) {
        mm_st4(reinterpret_cast<float4*>(g.mc_param) + i, pi[u]);
      } else {
#pragma unroll
        for (int k = 0; k < NP; ++k)
          if (k < world && (!(grads_zero & 4) || k == c.rank)) reinterpret_cast<float4*>(g.param[k])[i] = pi[u];
      }
    }
  }
}

template <int MODE, int WORLD, int U>
int launch_ddp_adam(const cnb_p2p_comm* comm, const cnb_p2p_group* group, float* exp_avg, float* exp_avg_sq, int64_t lo4, int64_t hi4, const AdamArgs& a,
                    int gz, cudaStream_t stream, const float* dev_scalars = nullptr) {
  int64_t blocks = (hi4 - lo4 + 256 * U - 1) / (256 * U);
  const int64_t cap = (int64_t)cnb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  k_ddp_adam<MODE, WORLD, U><<<(int)blocks, 256, 0, stream>>>(*comm, *group, reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), lo4,
                                                                hi4, a, gz, dev_scalars);
  return cnb_check_launch("ddp_adam_update");
}

int check_comm(const cnb_p2p_comm* c, const char* what) {
  CNB_REQUIRE(c != nullptr, "%s: null communicator", what);
  CNB_REQUIRE(c->world >= 1 && c->world <= CNB_MAX_PEERS && c->rank >= 0 && c->rank < c->world, "%s: bad world %d / rank %d", what, c->world, c->rank);
  CNB_REQUIRE(c->state != nullptr, "%s: null state", what);
  CNB_REQUIRE(c->channel >= 0 && c->channel < 4, "%s: channel %d outside 0..3", what, c->channel);
  for (int k = 0; k < c->world; ++k) CNB_REQUIRE(c->flags[k] != nullptr, "%s: null flag block of rank %d", what, k);
  return CNB_OK;
}

}  // namespace

extern "C" void cnb_p2p_owned_range(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi) {
  // slices are whole float4s: [rank*q, (rank+1)*q) with q = ceil(n/4 / world), clipped
  const int64_t n4 = n / 4, q = (n4 + world - 1) / world;
  int64_t a = (int64_t)rank * q, b = a + q;
  if (a > n4) a = n4;
  if (b > n4) b = n4;
  *lo = 4 * a;
  *hi = 4 * b;
}

extern "C" int cnb_p2p_barrier(const cnb_p2p_comm* comm, cnb_stream_t stream) {
  int rc = check_comm(comm, "p2p_barrier");
  if (rc) return rc;
  const uint64_t timeout_ns = comm->timeout_ms > 0 ? (uint64_t)comm->timeout_ms * 1000000ull : 10000000000ull;
  k_p2p_barrier<<<1, 32, 0, stream>>>(*comm, timeout_ns);
  return cnb_check_launch("p2p_barrier");
}

static int ddp_adam_update_impl(const cnb_p2p_comm* comm, const cnb_p2p_group* group, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, int32_t step, float inv_grad_scale, int32_t flags, const float* dev_scalars,
                                cnb_stream_t stream) {
  int rc = check_comm(comm, "ddp_adam_update");
  if (rc) return rc;
  CNB_REQUIRE(group && exp_avg && exp_avg_sq, "ddp_adam_update: null pointer");
  CNB_REQUIRE(n >= 0 && n % 4 == 0 && step >= 1, "ddp_adam_update: n must be a multiple of 4 (flat groups are), step >= 1");
  const bool multimem = (flags & CNB_P2P_MULTIMEM) != 0;
  for (int k = 0; k < comm->world; ++k) {
    CNB_REQUIRE(group->grad[k] && group->param[k], "ddp_adam_update: null peer buffer of rank %d", k);
    CNB_REQUIRE((((uintptr_t)group->grad[k] | (uintptr_t)group->param[k]) & 15) == 0, "ddp_adam_update: peer buffers must be 16-byte aligned");
  }
  CNB_REQUIRE((((uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "ddp_adam_update: moments must be 16-byte aligned");
  CNB_REQUIRE(!multimem || (group->mc_grad && group->mc_param), "ddp_adam_update: CNB_P2P_MULTIMEM needs the multicast mappings");
  int64_t lo, hi;
  cnb_p2p_owned_range(n, comm->rank, comm->world, &lo, &hi);
  if (hi <= lo) return CNB_OK;
  AdamArgs a;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.inv_scale = inv_grad_scale;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const int64_t lo4 = lo / 4, hi4 = hi / 4;
  const int gz = ((flags & CNB_P2P_GRADS_ZERO) ? 1 : 0) | ((flags & 16) ? 2 : 0) | ((flags & 32) ? 4 : 0);  // 16 / 32: timing aids (local loads only / local stores only)
  if (multimem) {
    if (flags & 0x100) return launch_ddp_adam<1, 0, 2>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);  // timing aids
    if (flags & 0x200) return launch_ddp_adam<1, 0, 4>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);
    return launch_ddp_adam<1, 0, 1>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);
  }
  switch (comm->world) {
    case 2: return launch_ddp_adam<0, 2, 1>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);
    case 4: return launch_ddp_adam<0, 4, 1>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);
    case 8: return launch_ddp_adam<0, 8, 1>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);
    default: return launch_ddp_adam<0, 0, 1>(comm, group, exp_avg, exp_avg_sq, lo4, hi4, a, gz, stream, dev_scalars);
  }
}

extern "C" int cnb_ddp_adam_update(const cnb_p2p_comm* comm, const cnb_p2p_group* group, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                   float beta1, float beta2, float eps, int32_t step, float inv_grad_scale, int32_t flags, cnb_stream_t stream) {
  return ddp_adam_update_impl(comm, group, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, inv_grad_scale, flags, nullptr, stream);
}

extern "C" int cnb_ddp_exchange_dev(const cnb_p2p_comm* comm, const cnb_p2p_group* group, float* exp_avg, float* exp_avg_sq, float* grad_own, int64_t n,
                                    const float* scalars, int32_t flags, int32_t channel, cnb_stream_t stream) {
  int rc = check_comm(comm, "ddp_exchange_dev");
  if (rc) return rc;
  CNB_REQUIRE(scalars != nullptr, "ddp_exchange_dev: null device scalars");
  CNB_REQUIRE(channel >= 0 && channel < 4, "ddp_exchange_dev: channel %d outside 0..3", channel);
  cnb_p2p_comm c = *comm;
  c.channel = channel;
  if ((rc = cnb_p2p_barrier(&c, stream))) return rc;                      // every rank's gradient of this group is final
  if ((rc = ddp_adam_update_impl(comm, group, exp_avg, exp_avg_sq, n, 1.0f, 0.9f, 0.999f, 1e-8f, 1, 1.0f, flags, scalars, stream))) return rc;
  if ((rc = cnb_p2p_barrier(&c, stream))) return rc;                      // every replica written, every rank's gradient consumed
  if (!(flags & CNB_P2P_GRADS_ZERO) && grad_own != nullptr && n > 0 &&
      cudaMemsetAsync(grad_own, 0, sizeof(float) * (size_t)n, stream) != cudaSuccess)
    return cnb_check_launch("ddp_exchange_dev clear");
  return CNB_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// The whole data-parallel optimiser step as ONE host call (what engine.Trainer._p2p_optimizer_step used to sequence from Python: ~0.3 ms of
// host time per step, which an end-to-end loop that reads the loss every step cannot hide):
//   stream:  barrier(ch 0)  ->  [groups with deferred == 0: update]  ->  barrier(ch 0)  ->  clear their gradients
//   side  :  (waits for the first barrier)  [groups with deferred != 0: update]  ->  barrier(ch 1)  ->  clear their gradients  ->  fence event
// cnb_ddp_wait_deferred(stream) makes a stream wait for the fence of the last step on this device (the consumer of the deferred groups'
// parameters -- the next step's field forward -- calls it; a no-op when nothing is in flight).
namespace {
struct DdpSide { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, fence = nullptr; bool pending = false, failed = false; };
DdpSide g_ddp[64];

DdpSide* ddp_side() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DdpSide& d = g_ddp[dev];
  if (d.failed) return nullptr;
  if (d.side == nullptr) {
    if (cudaStreamCreateWithFlags(&d.side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.fence, cudaEventDisableTiming) != cudaSuccess) {
      d.failed = true; d.side = nullptr;
      (void)cudaGetLastError();
      return nullptr;
    }
  }
  return &d;
}

int ddp_update_one(const cnb_p2p_comm* comm, const cnb_ddp_group_step& g, cudaStream_t st) {
  return cnb_ddp_adam_update(comm, g.group, g.exp_avg, g.exp_avg_sq, g.n, g.lr, g.beta1, g.beta2, g.eps, g.step, g.inv_grad_scale, g.flags, st);
}
int ddp_clear_one(const cnb_ddp_group_step& g, cudaStream_t st) {
  if ((g.flags & CNB_P2P_GRADS_ZERO) || g.grad_own == nullptr || g.n == 0) return CNB_OK;  // nobody wrote it this step
  if (cudaMemsetAsync(g.grad_own, 0, sizeof(float) * (size_t)g.n, st) != cudaSuccess) return cnb_check_launch("ddp_optimizer_step clear");
  return CNB_OK;
}
}  // namespace

extern "C" int cnb_ddp_optimizer_step(const cnb_p2p_comm* comm, const cnb_ddp_group_step* groups, int32_t n_groups, cnb_stream_t stream) {
  int rc = check_comm(comm, "ddp_optimizer_step");
  if (rc) return rc;
  CNB_REQUIRE(n_groups >= 0 && (n_groups == 0 || groups != nullptr), "ddp_optimizer_step: null groups");
  bool any_deferred = false, any_direct = false;
  for (int i = 0; i < n_groups; ++i) any_deferred = any_deferred || groups[i].deferred != 0;
  cnb_p2p_comm c0 = *comm, c1 = *comm;
  c0.channel = 0; c1.channel = 1;
  DdpSide* sd = any_deferred ? ddp_side() : nullptr;
  if (any_deferred && sd == nullptr) any_deferred = false;  // no side stream: everything on the caller's stream
  for (int i = 0; i < n_groups; ++i) any_direct = any_direct || !(any_deferred && groups[i].deferred);
  if (any_deferred) {
    // the deferred groups' exchange is self-contained on the side stream (its own barrier channel): it starts when the work enqueued on
    // `stream` so far -- this step's backward -- has finished, and adds nothing to `stream`
    if (cudaEventRecord(sd->fork, stream) != cudaSuccess || cudaStreamWaitEvent(sd->side, sd->fork, 0) != cudaSuccess) return cnb_check_launch("ddp_optimizer_step fork");
    if ((rc = cnb_p2p_barrier(&c1, sd->side))) return rc;  // every rank has finished its backward: the gradients are final
    for (int i = 0; i < n_groups; ++i)
      if (groups[i].deferred && (rc = ddp_update_one(comm, groups[i], sd->side))) return rc;
    if ((rc = cnb_p2p_barrier(&c1, sd->side))) return rc;  // every replica of those groups written, every rank's gradient consumed
    for (int i = 0; i < n_groups; ++i)
      if (groups[i].deferred && (rc = ddp_clear_one(groups[i], sd->side))) return rc;
    if (cudaEventRecord(sd->fence, sd->side) != cudaSuccess) return cnb_check_launch("ddp_optimizer_step fence");
    sd->pending = true;
  }
  if (!any_direct) return CNB_OK;
  if ((rc = cnb_p2p_barrier(&c0, stream))) return rc;
  for (int i = 0; i < n_groups; ++i)
    if (!(any_deferred && groups[i].deferred) && (rc = ddp_update_one(comm, groups[i], stream))) return rc;
  if ((rc = cnb_p2p_barrier(&c0, stream))) return rc;
  for (int i = 0; i < n_groups; ++i)
    if (!(any_deferred && groups[i].deferred) && (rc = ddp_clear_one(groups[i], stream))) return rc;
  return CNB_OK;
}

extern "C" int cnb_ddp_wait_deferred(cnb_stream_t stream) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return CNB_OK;
  DdpSide& d = g_ddp[dev];
  if (!d.pending) return CNB_OK;
  // the flag is not cleared: any number of streams may wait for the same fence; waiting for an already completed event costs nothing
  if (cudaStreamWaitEvent(stream, d.fence, 0) != cudaSuccess) return cnb_check_launch("ddp_wait_deferred");
  return CNB_OK;
}
