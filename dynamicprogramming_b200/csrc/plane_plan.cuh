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
// Two passes: plane_cells_kernel (one CTA per state-plane, fully parallel, memory-bound) finds the cells and writes
// the code words; plane_slots_kernel (one warp per chunk of `L` consecutive state-planes, sequential over the chunk
// because the slot state of step i depends on step i-1) assigns slots and writes the load lists.
#pragma once

#include <cuda_runtime.h>

namespace pi {

constexpr int kPlanMaxCells = 8;
constexpr int kPlanMaxOC = 16;
constexpr int kPlanMaxLoads = 80;
constexpr int kPlanMaxSlots = 80;
constexpr int kPlanHash = 32;
constexpr int kPlanFallback = 0x40000000;
constexpr unsigned short kPlanBadSlot = 0xffff;   // cs entry of a cell whose planes found no slot: its states gather from global memory
constexpr int kPlanThreads = 256;

struct PlaneRec {   // == PsRec in plane_sweep_src.cuh
    unsigned char n_early, n_late, n_cells, flags;
    unsigned int pad[3];
    unsigned short cs[kPlanMaxCells][kPlanMaxOC];   // slot byte offset / 16 of (cell, outer corner)
    unsigned int loads[kPlanMaxLoads];              // (slot << 24) | V-plane number; early loads first
};
static_assert(sizeof(PlaneRec) == 16 + 2 * kPlanMaxCells * kPlanMaxOC + 4 * kPlanMaxLoads, "PlaneRec layout");

// Per state-plane: its distinct successor cells (lower-corner V-plane numbers) and how many states fall into each.
struct PlaneCells {
    int n_cells;                             // <= kPlanMaxCells
    int n_unstaged;                          // live states whose cell did not fit (more than kPlanMaxCells distinct cells)
    int vp[kPlanMaxCells];
    unsigned short cnt[kPlanMaxCells];
    int pad[2];
};
static_assert(sizeof(PlaneCells) == 64, "PlaneCells layout");

struct PlanParams {
    const unsigned char* rows;   // the policy's rows; plane 0 = [base, f0, f1, f2] per state (16 B)
    unsigned char* prow0;        // out: [code, f0, f1, f2]
    PlaneRec* plan;              // out: one record per local state-plane
    PlaneCells* cells;           // scratch: one per local state-plane (written by plane_cells_kernel)
    unsigned long long* stats;   // out (atomic): [0] plane loads, [1] late loads, [2] fallback states, [3] live states,
                                 //               [4] staged cells, [5] state-planes
    int n_planes;                // local state-planes
    int L;                       // state-planes per chunk
    int P;                       // states per plane
    int NS;                      // shared-memory slots of the sweep
    int noc;                     // outer corners per cell (power of two)
    int pitch;                   // floats between slots in shared memory (>= P, multiple of 4)
    int ooff[kPlanMaxOC];        // V-plane offset of outer corner j
};

// Pass 1 — fully parallel, one CTA per state-plane: the distinct successor cells of the plane and the code word of every
// state (cell number << 16 | in-plane offset of the lower corner).  Memory-bound: reads and rewrites the first row plane.
constexpr int kCellThreads = 128;
__global__ void __launch_bounds__(kCellThreads) plane_cells_kernel(const PlanParams q) {
    __shared__ int htab[kPlanHash];             // distinct successor cells of the plane (open addressing)
    __shared__ int hcnt[kPlanHash];
    __shared__ int hcell[kPlanHash];            // hash entry -> cell number (or -1: does not fit)
    const int tid = threadIdx.x, lane = tid & 31;
    const int pl = blockIdx.x;
    const long long s0 = (long long)pl * q.P;
    if (tid < kPlanHash) { htab[tid] = -1; hcnt[tid] = 0; hcell[tid] = -1; }
    __syncthreads();
    constexpr int kPer = 8;   // states per thread (P <= 1024)
    uint4 w[kPer];
    int hs[kPer];
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int t = tid + r * kCellThreads;
        hs[r] = -1;
        w[r] = make_uint4(0xfffffffeu, 0u, 0u, 0u);
        if (t < q.P) w[r] = *reinterpret_cast<const uint4*>(q.rows + (size_t)(s0 + t) * 16u);
    }
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int t = tid + r * kCellThreads;
        const int base = (int)w[r].x;
        const bool live = t < q.P && base >= 0;
        const int vp = live ? base / q.P : -1;
        // one insertion per distinct cell per warp: the lanes of a warp mostly share 1-3 cells
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned same = __match_any_sync(act, vp);
            int h = -1;
            if ((__ffs(same) - 1) == lane) {
                h = (int)(((unsigned)vp * 2654435761u) >> 27);
                bool ok = false;
                for (int probe = 0; probe < kPlanHash && !ok; ++probe) {
                    const int old = atomicCAS(&htab[h], -1, vp);
                    ok = old == -1 || old == vp;
                    if (!ok) h = (h + 1) & (kPlanHash - 1);
                }
                if (ok) atomicAdd(&hcnt[h], __popc(same));
                else h = -1;
            }
            hs[r] = __shfl_sync(same, h, __ffs(same) - 1);
        }
    }
    __syncthreads();
    if (tid < 32) {   // number the occupied entries; the first kPlanMaxCells become the plane's cells
        const bool occ = htab[lane] != -1;
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        const int k = __popc(m & ((1u << lane) - 1u));
        PlaneCells* pc = q.cells + pl;
        if (occ && k < kPlanMaxCells) {
            hcell[lane] = k;
            pc->vp[k] = htab[lane];
            pc->cnt[k] = (unsigned short)min(hcnt[lane], 65535);
        }
        const int K = min(__popc(m), kPlanMaxCells);
        if (lane >= K && lane < kPlanMaxCells) { pc->vp[lane] = -1; pc->cnt[lane] = 0; }
        if (lane == 0) pc->n_cells = K;
    }
    __syncthreads();
    int n_unstaged = 0;
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int t = tid + r * kCellThreads;
        if (t < q.P) {
            int code = (int)w[r].x;
            if (code >= 0) {
                const int k = hs[r] >= 0 ? hcell[hs[r]] : -1;
                if (k >= 0) code = (k << 16) | (code - (code / q.P) * q.P);
                else { code = kPlanFallback; ++n_unstaged; }
            }
            *reinterpret_cast<uint4*>(q.prow0 + (size_t)(s0 + t) * 16u) = make_uint4((unsigned)code, w[r].y, w[r].z, w[r].w);
        }
    }
    // states outside every cell (rare: more than kPlanMaxCells distinct cells in a plane)
    const unsigned any = __ballot_sync(0xffffffffu, n_unstaged > 0);
    __shared__ int s_un;
    if (tid == 0) s_un = 0;
    __syncthreads();
    if (any && n_unstaged) atomicAdd(&s_un, n_unstaged);
    __syncthreads();
    if (tid == 0) q.cells[pl].n_unstaged = s_un;
}

// Pass 2 — one WARP per chunk of L consecutive state-planes, sequential over the chunk (the slot state of step i depends on
// step i-1): which slot holds which V-plane, which planes must be loaded before each step and whether that load can be
// issued one step ahead.  Works on the few cells per plane pass 1 found, not on states.
constexpr int kSlotWarps = 4;
__global__ void __launch_bounds__(kSlotWarps * 32) plane_slots_kernel(const PlanParams q) {
    __shared__ int s_slot_vp[kSlotWarps][kPlanMaxSlots];     // V-plane held by each slot (-1: none)
    __shared__ int s_slot_used[kSlotWarps][kPlanMaxSlots];   // last step that reads the slot
    __shared__ int s_cs[kSlotWarps][kPlanMaxCells * kPlanMaxOC];
    __shared__ unsigned s_early[kSlotWarps][kPlanMaxLoads], s_late[kSlotWarps][kPlanMaxLoads];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int chunk = blockIdx.x * kSlotWarps + wq;
    const int pl0 = chunk * q.L;
    if (pl0 >= q.n_planes) return;
    const int Lc = min(q.L, q.n_planes - pl0);
    int* slot_vp = s_slot_vp[wq];
    int* slot_used = s_slot_used[wq];
    int* cs = s_cs[wq];
    unsigned* early = s_early[wq];
    unsigned* late = s_late[wq];
    for (int s = lane; s < kPlanMaxSlots; s += 32) { slot_vp[s] = -1; slot_used[s] = -1000; }
    __syncwarp();
    int noc_shift = 0;
    while ((1 << noc_shift) < q.noc) ++noc_shift;
    unsigned long long a_loads = 0, a_late = 0, a_fb = 0, a_live = 0, a_cells = 0;

    for (int i = 0; i < Lc; ++i) {
        const PlaneCells pc = q.cells[pl0 + i];
        const int K = pc.n_cells;
        const int T = K << noc_shift;                       // (cell, outer corner) pairs of this plane
        // resident corner planes keep their slot
        for (int m = lane; m < T; m += 32) {
            const int vp = pc.vp[m >> noc_shift] + q.ooff[m & (q.noc - 1)];
            int found = -1;
            for (int s = 0; s < q.NS; ++s)
                if (slot_vp[s] == vp) found = s;
            cs[m] = found;
            if (found >= 0) slot_used[found] = i;
        }
        __syncwarp();
        // allocate slots for the missing planes, one at a time (slot scan in parallel across the lanes)
        int n_early = 0, n_late = 0;
        unsigned bad = 0;                                   // cells that could not be staged (no slot left)
        for (int m0 = 0; m0 < T; m0 += 32) {
            const int mine_m = m0 + lane;
            unsigned todo = __ballot_sync(0xffffffffu, mine_m < T && cs[mine_m] < 0);
            while (todo) {
                const int m = m0 + __ffs(todo) - 1;
                todo &= todo - 1;
                const int k = m >> noc_shift;
                if ((bad >> k) & 1u) continue;
                const int vp = pc.vp[k] + q.ooff[m & (q.noc - 1)];
                int mine = -1, best_early = -1, best_late = -1;
                for (int s = lane; s < q.NS; s += 32) {
                    const int u = slot_used[s];
                    if (slot_vp[s] == vp && u == i) mine = s;      // allocated earlier in this step (another cell shares the plane)
                    if (u < i - 1 && best_early < 0) best_early = s;
                    if (u == i - 1 && best_late < 0) best_late = s;
                }
                const unsigned mm = __ballot_sync(0xffffffffu, mine >= 0);
                const unsigned me = __ballot_sync(0xffffffffu, best_early >= 0);
                const unsigned ml = __ballot_sync(0xffffffffu, best_late >= 0);
                int slot = -1, kind = 0;   // kind 1: early, 2: late
                if (mm) slot = __shfl_sync(0xffffffffu, mine, __ffs(mm) - 1);
                else if (me) { slot = __shfl_sync(0xffffffffu, best_early, __ffs(me) - 1); kind = i > 0 ? 1 : 2; }   // first plane of a chunk: nothing was issued ahead
                else if (ml) { slot = __shfl_sync(0xffffffffu, best_late, __ffs(ml) - 1); kind = 2; }
                if (slot < 0 || (kind && n_early + n_late >= kPlanMaxLoads)) {
                    bad |= 1u << k;
                } else {
                    if (lane == 0) {
                        cs[m] = slot;
                        if (kind) {
                            slot_vp[slot] = vp;
                            slot_used[slot] = i;
                            const unsigned e = ((unsigned)slot << 24) | (unsigned)vp;
                            if (kind == 1) early[n_early] = e;
                            else late[n_late] = e;
                        }
                    }
                    if (kind == 1) ++n_early;
                    else if (kind == 2) ++n_late;
                }
                __syncwarp();
            }
        }
        // outputs
        PlaneRec* rec = q.plan + pl0 + i;
        if (lane == 0) {
            rec->n_early = (unsigned char)n_early;
            rec->n_late = (unsigned char)n_late;
            rec->n_cells = (unsigned char)K;
            rec->flags = (unsigned char)bad;
            rec->pad[0] = rec->pad[1] = rec->pad[2] = 0u;
        }
        for (int t = lane; t < kPlanMaxCells * kPlanMaxOC; t += 32) {
            const int k = t / kPlanMaxOC, j = t - k * kPlanMaxOC;
            unsigned short v = 0;
            if (k < K && j < q.noc) {
                if ((bad >> k) & 1u) v = kPlanBadSlot;
                else v = (unsigned short)(((size_t)cs[(k << noc_shift) + j] * (size_t)q.pitch * 4u) >> 4);
            }
            rec->cs[k][j] = v;
        }
        for (int t = lane; t < n_early + n_late; t += 32) rec->loads[t] = t < n_early ? early[t] : late[t - n_early];
        a_loads += (unsigned long long)(n_early + n_late);
        a_late += (unsigned long long)n_late;
        a_fb += (unsigned long long)pc.n_unstaged;
        a_live += (unsigned long long)pc.n_unstaged;
        for (int k = 0; k < K; ++k) {
            a_live += pc.cnt[k];
            if ((bad >> k) & 1u) a_fb += pc.cnt[k];
            else a_cells += 1;
        }
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(q.stats + 0, a_loads);
        atomicAdd(q.stats + 1, a_late);
        atomicAdd(q.stats + 2, a_fb);
        atomicAdd(q.stats + 3, a_live);
        atomicAdd(q.stats + 4, a_cells);
        atomicAdd(q.stats + 5, (unsigned long long)Lc);
    }
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
