// Gradient exchange of the data-parallel path as ONE kernel over NVSwitch multicast memory (NVLS): the mean over ranks of the
// flat gradient buffer the backward kernels accumulated into (the reference gets it from DDP's bucketed NCCL all-reduce,
// fruit_nerf/fruit_pipeline.py:116-118).
//
// The buffer lives in symmetric memory: every rank maps its own copy (unicast) and the multicast object that spans all copies.
// Rank r owns the slice [r n/W, (r+1) n/W):
//   multimem.ld_reduce.add   one load returns the SUM of that address over all W copies -- the reduction happens in the switch,
//                            W-1 of the W operands never cross this GPU's links as separate transfers;
//   x 1/W                    (DDP averages);
//   multimem.st              one store writes the mean into all W copies (switch-side broadcast).
// Per GPU and direction that is ~n bytes on the wire for an n-byte buffer -- the all-reduce lower bound for in-switch reduction --
// in one launch, with no staging copies.  Ranks meet twice per launch on flags in each other's signal pads (symmetric too):
// block b of every rank waits for block b of every peer to have started (their gradients are complete: stream order) and, at the
// end, to have finished its stores (nobody reads a partially written buffer or re-zeroes one that is still being read).
// `wire_bf16`: the operands cross the wire as bf16 (half the bytes; accumulation in fp32 inside the switch): each rank first rounds
// its fp32 gradients into a bf16 staging half of the symmetric region, the switch reduces / broadcasts bf16, and each rank widens
// the result back -- three phases, still one launch (a grid-wide barrier separates them locally).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "fnr_common.cuh"
#include "fnr_kernels.h"

namespace fnr {
namespace {

constexpr int kThreads = 512;
constexpr int kUnroll = 8;  // 128 B of multimem loads in flight per thread: a slice is a handful of iterations, not dozens of switch round trips

struct NvlsArgs {
  float* mc;             // multicast address of the fp32 region
  float* local;          // this rank's mapping of the fp32 region
  __nv_bfloat16* mc16;   // multicast / local address of the bf16 staging region (wire_bf16 only)
  __nv_bfloat16* local16;
  uint32_t* const* pads; // device array [world]: signal pad of every rank (own pad at [rank])
  unsigned int* grid_counter;  // local grid barrier (wire_bf16 only): zero before the first launch, self-resetting
  long long n;           // fp32 elements, multiple of 4 * world
  int rank, world, slot_base;
  float scale;
};

// A rank that never launches (crashed peer, mismatched call sequence) must not hang the GPU: after ~4 s of spinning the kernel traps,
// which surfaces as a CUDA error on the stream instead of a dead device.
constexpr long long kSpinLimitCycles = 8000000000ll;

__device__ __forceinline__ void put_signal(uint32_t* addr) {
  // release at system scope: everything this block wrote (multimem.st included) is visible before the peer sees the flag
  __threadfence_system();
  const long long t0 = clock64();
  while (atomicCAS_system(addr, 0u, 1u) != 0u) {
    if (clock64() - t0 > kSpinLimitCycles) __trap();
  }
}
__device__ __forceinline__ void wait_signal(uint32_t* addr) {
  const long long t0 = clock64();
  while (atomicCAS_system(addr, 1u, 0u) != 1u) {
    if (clock64() - t0 > kSpinLimitCycles) __trap();
  }
  __threadfence_system();
}

// block b of every rank meets block b of every other rank
__device__ __forceinline__ void cross_rank_barrier(const NvlsArgs& a) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < a.world && t != a.rank) {
    const int slot = a.slot_base + blockIdx.x * a.world;
    put_signal(a.pads[t] + slot + a.rank);
    wait_signal(a.pads[a.rank] + slot + t);
  }
  __syncthreads();
}

// all blocks of THIS rank (grid <= number of SMs, one block per SM: every block is resident)
__device__ __forceinline__ void grid_barrier(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    const unsigned int target = (ticket / gridDim.x + 1u) * gridDim.x;  // the counter keeps counting across launches
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mm_ld_reduce_f32(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f32(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 8 bf16 values: summed over ranks with fp32 accumulation inside the switch, returned as 4 packed bf16x2
__device__ __forceinline__ uint4 mm_ld_reduce_bf16(const __nv_bfloat16* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_bf16(__nv_bfloat16* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t v, float s) {
  const float lo = __uint_as_float(v << 16) * s, hi = __uint_as_float(v & 0xffff0000u) * s;
  __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&r);
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 1) nvls_allreduce_mean_kernel(const NvlsArgs a) {
  const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x, nthreads = (long long)gridDim.x * kThreads;
  if constexpr (kBf16) {
    // phase 0: round this rank's gradients into its bf16 staging copy (8 elements per thread and step)
    for (long long i = tid * 8; i < a.n; i += nthreads * 8) {
      const float4 u = *reinterpret_cast<const float4*>(a.local + i), w = *reinterpret_cast<const float4*>(a.local + i + 4);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(u.x, u.y), p1 = __floats2bfloat162_rn(u.z, u.w), p2 = __floats2bfloat162_rn(w.x, w.y),
                     p3 = __floats2bfloat162_rn(w.z, w.w);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&p0);
      o.y = *reinterpret_cast<uint32_t*>(&p1);
      o.z = *reinterpret_cast<uint32_t*>(&p2);
      o.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(a.local16 + i) = o;
    }
    grid_barrier(a.grid_counter);  // peers read what ANY of this rank's blocks wrote
  }
  cross_rank_barrier(a);

  const long long per = a.n / a.world, lo = per * a.rank;
  if constexpr (!kBf16) {
    const long long vecs = per / 4;
    for (long long v0 = tid; v0 < vecs; v0 += nthreads * kUnroll) {
      float4 r[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long v = v0 + u * nthreads;
        if (v < vecs) r[u] = mm_ld_reduce_f32(a.mc + lo + 4 * v);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long v = v0 + u * nthreads;
        if (v < vecs) mm_st_f32(a.mc + lo + 4 * v, make_float4(r[u].x * a.scale, r[u].y * a.scale, r[u].z * a.scale, r[u].w * a.scale));
      }
    }
  } else {
    const long long vecs = per / 8;
    for (long long v0 = tid; v0 < vecs; v0 += nthreads * kUnroll) {
      uint4 r[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long v = v0 + u * nthreads;
        if (v < vecs) r[u] = mm_ld_reduce_bf16(a.mc16 + lo + 8 * v);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long v = v0 + u * nthreads;
        if (v < vecs)
          mm_st_bf16(a.mc16 + lo + 8 * v, make_uint4(scale_bf16x2(r[u].x, a.scale), scale_bf16x2(r[u].y, a.scale), scale_bf16x2(r[u].z, a.scale),
                                                      scale_bf16x2(r[u].w, a.scale)));
      }
    }
  }
  cross_rank_barrier(a);

  if constexpr (kBf16) {
    grid_barrier(a.grid_counter);  // every block of this rank has seen every peer finish its stores
    for (long long i = tid * 8; i < a.n; i += nthreads * 8) {
      const uint4 o = *reinterpret_cast<const uint4*>(a.local16 + i);
      const float4 u = make_float4(__uint_as_float(o.x << 16), __uint_as_float(o.x & 0xffff0000u), __uint_as_float(o.y << 16), __uint_as_float(o.y & 0xffff0000u));
      const float4 w = make_float4(__uint_as_float(o.z << 16), __uint_as_float(o.z & 0xffff0000u), __uint_as_float(o.w << 16), __uint_as_float(o.w & 0xffff0000u));
      *reinterpret_cast<float4*>(a.local + i) = u;
      *reinterpret_cast<float4*>(a.local + i + 4) = w;
    }
  }
}

}  // namespace
}  // namespace fnr

using namespace fnr;

extern "C" int fnr_nvls_allreduce_mean(const fnr_nvls_desc* d, size_t numel, int32_t wire_bf16, void* stream) {
  if (!d || !d->multicast_ptr || !d->local_ptr || !d->signal_pads || d->world_size < 2 || d->rank < 0 || d->rank >= d->world_size) {
    set_error("invalid NVLS descriptor");
    return FNR_ERR_INVALID_ARGUMENT;
  }
  if (numel == 0) return FNR_OK;
  if (numel % (size_t)(8 * d->world_size) != 0) {
    set_error("numel (%zu) must be a multiple of 8 * world_size (pad the flat buffer)", numel);
    return FNR_ERR_INVALID_ARGUMENT;
  }
  if (wire_bf16 && (!d->multicast_bf16 || !d->local_bf16 || !d->grid_counter)) {
    set_error("wire_bf16 needs the bf16 staging region and a grid counter");
    return FNR_ERR_INVALID_ARGUMENT;
  }
  const int avail = (d->signal_slots - d->signal_slot_base) / d->world_size;
  // one block per SM (grid barrier); measured at N = 2: 64 blocks x 4 vectors per thread ran 16 dependent load -> store iterations per
  // slice and reached 45 % of the link rate -- the exchange is latency-bound unless (almost) the whole slice is in flight at once
  int blocks = sm_count() < 144 ? sm_count() : 144;
  if (blocks > avail) blocks = avail;
  if (blocks < 1) {
    set_error("signal pad too small: %d slots for world size %d", d->signal_slots, d->world_size);
    return FNR_ERR_INVALID_ARGUMENT;
  }
  NvlsArgs a;
  a.mc = static_cast<float*>(d->multicast_ptr);
  a.local = static_cast<float*>(d->local_ptr);
  a.mc16 = static_cast<__nv_bfloat16*>(d->multicast_bf16);
  a.local16 = static_cast<__nv_bfloat16*>(d->local_bf16);
  a.pads = reinterpret_cast<uint32_t* const*>(d->signal_pads);
  a.grid_counter = static_cast<unsigned int*>(d->grid_counter);
  a.n = (long long)numel;
  a.rank = d->rank;
  a.world = d->world_size;
  a.slot_base = d->signal_slot_base;
  a.scale = 1.0f / (float)d->world_size;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (wire_bf16) nvls_allreduce_mean_kernel<true><<<blocks, kThreads, 0, st>>>(a);
  else nvls_allreduce_mean_kernel<false><<<blocks, kThreads, 0, st>>>(a);
  return check_launch("nvls_allreduce_mean_kernel");
}
