// Exclusive scan of u32 counters in tiles of 4096 (shared by the MSM digit sort, msm.cuh, and the device-side CSR
// build of the coefficient records, prove.cu): k_scan_tile scans each tile and leaves its total, k_scan_totals scans
// the totals, k_scan_add adds them back and writes a second copy (the scatter cursors).
#pragma once
#include "bigint.cuh"

namespace b200 {

static const int SCAN_TILE = 4096;

static __global__ void __launch_bounds__(1024) k_scan_tile(u32 *__restrict__ data, u32 *__restrict__ totals) {
    __shared__ u32 warp_sums[32];
    uint4 *p = reinterpret_cast<uint4 *>(data) + (size_t)blockIdx.x * 1024 + threadIdx.x;
    uint4 v = *p;
    u32 t0 = v.x, t1 = t0 + v.y, t2 = t1 + v.z, t3 = t2 + v.w;
    u32 incl = t3;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 ws = warp_sums[lane];
        u32 wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += y;
        }
        warp_sums[lane] = wi - ws;  // exclusive
        if (lane == 31) totals[blockIdx.x] = wi;
    }
    __syncthreads();
    u32 base = warp_sums[wid] + incl - t3;
    uint4 o;
    o.x = base; o.y = base + t0; o.z = base + t1; o.w = base + t2;
    *p = o;
}

static __global__ void __launch_bounds__(1024) k_scan_totals(u32 *__restrict__ totals, u32 ntiles) {
    __shared__ u32 warp_sums[32];
    __shared__ u32 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 base = 0; base < ntiles; base += 1024) {
        u32 i = base + threadIdx.x;
        u32 v = i < ntiles ? totals[i] : 0;
        u32 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            u32 ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += y;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        u32 carry = carry_s;
        u32 excl = carry + warp_sums[wid] + incl - v;
        if (i < ntiles) totals[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(1024) k_scan_add(u32 *__restrict__ data, const u32 *__restrict__ totals,
                                                   u32 *__restrict__ copy) {
    uint4 *p = reinterpret_cast<uint4 *>(data) + (size_t)blockIdx.x * 1024 + threadIdx.x;
    u32 add = totals[blockIdx.x];
    uint4 v = *p;
    v.x += add; v.y += add; v.z += add; v.w += add;
    *p = v;
    reinterpret_cast<uint4 *>(copy)[(size_t)blockIdx.x * 1024 + threadIdx.x] = v;  // scatter cursors
}


}  // namespace b200
