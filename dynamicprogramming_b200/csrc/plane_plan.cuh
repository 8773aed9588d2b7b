// plane_plan.cuh — per-policy plan of the plane-staged evaluation sweep (plane_sweep_src.cuh).
//
// The policy's rows are fixed for the thousands of sweeps of one policy evaluation
// (src/cuda_policy_iteration.py:300-336), so everything the sweep needs to know about WHICH V-planes
// a state-plane reads is computed once per policy, here, instead of once per sweep:
//
//   * for every state-plane (the PS_P states that share their outer-dimension indices) the distinct
//     successor cells of the outer dimensions (<= kPlanMaxCells) and, per state, the code word
//     (cell number << 16 | in-plane offset of the lower corner) that replaces the row's base index;
//   * a shared-memory slot for each of the 2^(D-2) corner V-planes of every cell, assigned while
//     walking the chunk's state-planes in order: a V-plane already resident from the previous
//     state-plane keeps its slot (no load), a new one takes an idle slot — "early" when the slot was
//     not read by the previous state-plane either (the load overlaps that plane's backups), "late"
//     when it only becomes free once the previous plane is done;
//   * cells that do not fit (more than kPlanMaxCells per plane, no slot left) are flagged: their states
//     gather from global memory in the sweep.
//
// One CTA per chunk of `L` consecutive state-planes, sequential over the chunk (the slot state of
// step i depends on step i-1), parallel over the plane's states and over slot searches.
#pragma once

#include <cuda_runtime.h>

namespace pi {

constexpr int kPlanMaxCells = 8;
constexpr int kPlanMaxOC = 16;
constexpr int kPlanMaxLoads = 80;
constexpr int kPlanMaxSlots = 80;
constexpr int kPlanHash = 32;
constexpr int kPlanFallback = 0x40000000;
constexpr int kPlanThreads = 256;

struct PlaneRec {   // == PsRec in plane_sweep_src.cuh
    unsigned char n_early, n_late, n_cells, flags;
    unsigned int pad[3];
    unsigned short cs[kPlanMaxCells][kPlanMaxOC];   // slot byte offset / 16 of (cell, outer corner)
    unsigned int loads[kPlanMaxLoads];              // (slot << 24) | V-plane number; early loads first
};
static_assert(sizeof(PlaneRec) == 16 + 2 * kPlanMaxCells * kPlanMaxOC + 4 * kPlanMaxLoads, "PlaneRec layout");

struct PlanParams {
    const unsigned char* rows;   // the policy's rows; plane 0 = [base, f0, f1, f2] per state (16 B)
    unsigned char* prow0;        // out: [code, f0, f1, f2]
    PlaneRec* plan;              // out: one record per local state-plane
    unsigned long long* stats;   // out (atomic): [0] plane loads, [1] late loads, [2] fallback states, [3] live states,
                                 //               [4] staged cells, [5] state-planes
    long long plane0;            // global number of this rank's first state-plane
    int n_planes;                // local state-planes
    int L;                       // state-planes per chunk
    int P;                       // states per plane
    int NS;                      // shared-memory slots of the sweep
    int noc;                     // outer corners per cell
    int pitch;                   // floats between slots in shared memory (>= P, multiple of 4)
    int ooff[kPlanMaxOC];        // V-plane offset of outer corner j
};

__global__ void __launch_bounds__(kPlanThreads) plane_plan_kernel(const PlanParams q) {
    __shared__ int slot_vp[kPlanMaxSlots];      // V-plane held by each slot (-1: none)
    __shared__ int slot_used[kPlanMaxSlots];    // last step that reads the slot
    __shared__ int htab[kPlanHash];             // distinct successor cells of the plane (open addressing)
    __shared__ int hcell[kPlanHash];            // hash entry -> cell number (or -1: not staged)
    __shared__ int cells[kPlanMaxCells];
    __shared__ int cs[kPlanMaxCells][kPlanMaxOC];
    __shared__ int cell_bad[kPlanMaxCells];
    __shared__ unsigned early[kPlanMaxLoads], late[kPlanMaxLoads];
    __shared__ int n_cells, n_early, n_late;
    __shared__ unsigned long long acc[6];

    const int tid = threadIdx.x, lane = tid & 31;
    const int pl0 = blockIdx.x * q.L;
    const int Lc = min(q.L, q.n_planes - pl0);
    for (int s = tid; s < kPlanMaxSlots; s += kPlanThreads) { slot_vp[s] = -1; slot_used[s] = -1000; }
    if (tid < 6) acc[tid] = 0ull;
    __syncthreads();

    constexpr int kPer = 4;   // states per thread (P <= 1024)
    for (int i = 0; i < Lc; ++i) {
        const long long s0 = (long long)(pl0 + i) * q.P;
        if (tid < kPlanHash) { htab[tid] = -1; hcell[tid] = -1; }
        if (tid < kPlanMaxCells) { cell_bad[tid] = 0; cells[tid] = 0; }
        if (tid == 0) { n_cells = 0; n_early = 0; n_late = 0; }
        __syncthreads();

        // 1. distinct successor cells (lower-corner V-plane = base / P)
        int hs[kPer], bases[kPer];
#pragma unroll
        for (int r = 0; r < kPer; ++r) {
            const int t = tid + r * kPlanThreads;
            hs[r] = -1;
            bases[r] = -2;
            if (t < q.P) {
                const int base = *reinterpret_cast<const int*>(q.rows + (size_t)(s0 + t) * 16u);
                bases[r] = base;
                if (base >= 0) {
                    const int vp = base / q.P;
                    int h = (int)(((unsigned)vp * 2654435761u) >> 27);
                    for (int probe = 0; probe < kPlanHash; ++probe) {
                        const int old = atomicCAS(&htab[h], -1, vp);
                        if (old == -1 || old == vp) { hs[r] = h; break; }
                        h = (h + 1) & (kPlanHash - 1);
                    }
                }
            }
        }
        __syncthreads();
        // 2. number the occupied entries
        if (tid < 32) {
            const bool occ = htab[lane] != -1;
            const unsigned m = __ballot_sync(0xffffffffu, occ);
            const int k = __popc(m & ((1u << lane) - 1u));
            if (occ && k < kPlanMaxCells) { hcell[lane] = k; cells[k] = htab[lane]; }
            if (lane == 0) n_cells = min(__popc(m), kPlanMaxCells);
        }
        __syncthreads();
        const int K = n_cells;
        // 3. resident corner planes keep their slot
        if (tid < K * q.noc) {
            const int k = tid / q.noc, j = tid - k * q.noc;
            const int vp = cells[k] + q.ooff[j];
            int found = -1;
            for (int s = 0; s < q.NS; ++s)
                if (slot_vp[s] == vp) found = s;
            cs[k][j] = found;
            if (found >= 0) slot_used[found] = i;
        }
        __syncthreads();
        // 4. allocate slots for the missing planes (warp 0, one plane at a time, slot scan in parallel)
        if (tid < 32) {
            for (int m = 0; m < K * q.noc; ++m) {
                const int k = m / q.noc, j = m - k * q.noc;
                if (cs[k][j] >= 0 || cell_bad[k]) continue;   // uniform across the warp (shared memory)
                const int vp = cells[k] + q.ooff[j];
                // allocated earlier in this step (another cell shares the plane)?
                int mine = -1, best_early = -1, best_late = -1;
                for (int s = lane; s < q.NS; s += 32) {
                    if (slot_vp[s] == vp && slot_used[s] == i) mine = s;
                    if (slot_used[s] < i - 1 && best_early < 0) best_early = s;
                    if (slot_used[s] == i - 1 && best_late < 0) best_late = s;
                }
                const unsigned mm = __ballot_sync(0xffffffffu, mine >= 0);
                const unsigned me = __ballot_sync(0xffffffffu, best_early >= 0);
                const unsigned ml = __ballot_sync(0xffffffffu, best_late >= 0);
                int slot = -1, kind = 0;   // kind 1: early, 2: late
                if (mm) slot = __shfl_sync(0xffffffffu, mine, __ffs(mm) - 1);
                else if (me && i > 0) { slot = __shfl_sync(0xffffffffu, best_early, __ffs(me) - 1); kind = 1; }
                else if (me) { slot = __shfl_sync(0xffffffffu, best_early, __ffs(me) - 1); kind = 2; }   // first plane of a chunk: nothing was issued ahead
                else if (ml) { slot = __shfl_sync(0xffffffffu, best_late, __ffs(ml) - 1); kind = 2; }
                if (lane == 0) {
                    if (slot < 0 || (kind && n_early + n_late >= kPlanMaxLoads)) {
                        cell_bad[k] = 1;
                    } else {
                        cs[k][j] = slot;
                        if (kind) {
                            slot_vp[slot] = vp;
                            slot_used[slot] = i;
                            const unsigned e = ((unsigned)slot << 24) | (unsigned)vp;
                            if (kind == 1) early[n_early++] = e;
                            else late[n_late++] = e;
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // 5. outputs
        PlaneRec* rec = q.plan + pl0 + i;
        if (tid == 0) {
            rec->n_early = (unsigned char)n_early;
            rec->n_late = (unsigned char)n_late;
            rec->n_cells = (unsigned char)K;
            rec->flags = 0;
            rec->pad[0] = rec->pad[1] = rec->pad[2] = 0u;
            int staged = 0;
            for (int k = 0; k < K; ++k) staged += cell_bad[k] ? 0 : 1;
            acc[0] += (unsigned long long)(n_early + n_late);
            acc[1] += (unsigned long long)n_late;
            acc[4] += (unsigned long long)staged;
            acc[5] += 1ull;
        }
        if (tid < kPlanMaxCells * kPlanMaxOC) {
            const int k = tid / kPlanMaxOC, j = tid - k * kPlanMaxOC;
            int slot = 0;
            if (k < K && j < q.noc && !cell_bad[k]) slot = cs[k][j];
            rec->cs[k][j] = (unsigned short)(((size_t)slot * (size_t)q.pitch * 4u) >> 4);
        }
        for (int t = tid; t < n_early + n_late; t += kPlanThreads) rec->loads[t] = t < n_early ? early[t] : late[t - n_early];
        unsigned n_fb = 0, n_live = 0;
#pragma unroll
        for (int r = 0; r < kPer; ++r) {
            const int t = tid + r * kPlanThreads;
            if (t < q.P) {
                const uint4 v = *reinterpret_cast<const uint4*>(q.rows + (size_t)(s0 + t) * 16u);
                int code = bases[r];
                if (code >= 0) {
                    ++n_live;
                    const int k = hs[r] >= 0 ? hcell[hs[r]] : -1;
                    if (k >= 0 && !cell_bad[k]) code = (k << 16) | (code - (code / q.P) * q.P);
                    else { code = kPlanFallback; ++n_fb; }
                }
                *reinterpret_cast<uint4*>(q.prow0 + (size_t)(s0 + t) * 16u) = make_uint4((unsigned)code, v.y, v.z, v.w);
            }
        }
        if (n_fb) atomicAdd(&acc[2], (unsigned long long)n_fb);
        if (n_live) atomicAdd(&acc[3], (unsigned long long)n_live);
        __syncthreads();
    }
    if (tid < 6 && acc[tid]) atomicAdd(q.stats + tid, acc[tid]);
}


// Layout probe: how many distinct successor cells of the outer dimensions does a state-plane of P states see
// (rows of ONE action, planes [0, n_planes) of a temporary table)?  1 when the dynamics are translation-invariant
// in the two fastest-stored dimensions.  out[0] += distinct cells, out[1] += planes with a live state.
__global__ void __launch_bounds__(kPlanThreads) plane_cells_probe_kernel(const unsigned char* rows, int P, int n_planes,
                                                                         unsigned long long* out) {
    __shared__ int htab[kPlanHash];
    __shared__ int overflow;
    const int tid = threadIdx.x;
    for (int pl = blockIdx.x; pl < n_planes; pl += gridDim.x) {
        if (tid < kPlanHash) htab[tid] = -1;
        if (tid == 0) overflow = 0;
        __syncthreads();
        for (int t = tid; t < P; t += kPlanThreads) {
            const int base = *reinterpret_cast<const int*>(rows + ((size_t)pl * P + t) * 16u);
            if (base < 0) continue;
            const int vp = base / P;
            int h = (int)(((unsigned)vp * 2654435761u) >> 27);
            bool ok = false;
            for (int probe = 0; probe < kPlanHash && !ok; ++probe) {
                const int old = atomicCAS(&htab[h], -1, vp);
                ok = old == -1 || old == vp;
                h = (h + 1) & (kPlanHash - 1);
            }
            if (!ok) overflow = 1;
        }
        __syncthreads();
        if (tid < 32) {
            const unsigned m = __ballot_sync(0xffffffffu, htab[tid] != -1);
            if (tid == 0 && m) {
                atomicAdd(out, (unsigned long long)(overflow ? 4 * kPlanHash : __popc(m)));
                atomicAdd(out + 1, 1ull);
            }
        }
        __syncthreads();
    }
}

}  // namespace pi
