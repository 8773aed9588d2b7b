// pi_kernels.cuh — ahead-of-time sm_100a kernels of the policy-iteration engine.
//
// Everything here works on *compact transition rows*: for one (state, action)
// pair a row is W = D + 2 32-bit words
//     [ base | frac_0 .. frac_{D-1} | reward ]
// where `base` is the flat index of the lower corner of the interpolation cell
// (or a PI_ROW_* sentinel) and frac_d the position inside the cell.  The 2^D
// corner indices and weights of the reference's get_barycentric_{2,4,6}d
// (src/cuda_policy_iteration.py:183-210, :580-614, :1007-1042) are rebuilt in
// registers with the reference's exact operation order, so results are
// bit-identical while a 6-D row costs 32 B of HBM instead of 516 B.
//
// Row storage ("plane-SoA"): the W words of a row are split into 16-byte,
// 8-byte and 4-byte planes; plane p of a table holding n_pad states is one
// contiguous array, so lane i of a warp reads/writes 16 B at stride 16 B:
// every row access is a fully coalesced 128-bit (or 64/32-bit) transaction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define PI_ROW_TERMINATED (-1)
#define PI_ROW_ABSORBING (-2)

namespace pi {

constexpr int kBlock = 256;
constexpr int kMaxDims = 6;

// ---------------------------------------------------------------------------
// Device control block: lets a stream of identical sweep launches sequence
// itself (sweep counter, ping-pong parity, residual, convergence flag) with no
// host round trip per sweep.  Mirrors the loop state of policy_evaluation()
// (src/cuda_policy_iteration.py:300-336).
// ---------------------------------------------------------------------------
struct Ctl {
    int base;                   // sweeps completed before the batch in flight
    int parity0;                // V buffer that was "current" when the evaluation started
    int done;                   // 1 once a sync-point residual was < theta
    int conv_sweep;             // index i of the sweep at which the evaluation converged
    float last_delta;           // residual of the last examined sweep (local to this rank)
    float check_delta;          // residual examined at the last sync point (global)
    unsigned long long changed; // improvement: states whose action changed (local)
    unsigned int pad[8];
};

struct GridDesc {
    int shape[kMaxDims];
    int stride[kMaxDims];
    float lo[kMaxDims];
    float hi[kMaxDims];
};

// ---------------------------------------------------------------------------
// Row layout helpers
// ---------------------------------------------------------------------------
template <int D>
struct Row {
    static constexpr int W = D + 2;
    static constexpr int N4 = W / 4;        // 16-byte planes
    static constexpr int N2 = (W % 4) / 2;  // 8-byte plane (0/1)
    static constexpr int N1 = W % 2;        // 4-byte plane (0/1)
    static constexpr int kBytes = W * 4;
};

__host__ __device__ inline size_t row_table_bytes(int D, long long n_pad) {
    return (size_t)(D + 2) * 4u * (size_t)n_pad;
}

// streaming (read-once) loads: do not pollute L1, mark evict-first in L2
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ld_stream4(const void* p) {
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream8(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, unsigned v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Load the row of local state `s` from the table that starts at `tab`.
template <int D>
__device__ __forceinline__ void load_row(const unsigned char* __restrict__ tab, long long n_pad, long long s,
                                         unsigned (&w)[Row<D>::W]) {
    using R = Row<D>;
#pragma unroll
    for (int p = 0; p < R::N4; ++p) {
        uint4 v = ld_stream16(tab + (size_t)p * 16u * (size_t)n_pad + (size_t)s * 16u);
        w[4 * p + 0] = v.x; w[4 * p + 1] = v.y; w[4 * p + 2] = v.z; w[4 * p + 3] = v.w;
    }
    if constexpr (R::N2 != 0) {
        uint2 v = ld_stream8(tab + (size_t)R::N4 * 16u * (size_t)n_pad + (size_t)s * 8u);
        w[4 * R::N4 + 0] = v.x; w[4 * R::N4 + 1] = v.y;
    }
    if constexpr (R::N1 != 0) {
        w[R::W - 1] = ld_stream4(tab + ((size_t)R::N4 * 16u + (size_t)R::N2 * 8u) * (size_t)n_pad + (size_t)s * 4u);
    }
}

template <int D>
__device__ __forceinline__ void store_row(unsigned char* __restrict__ tab, long long n_pad, long long s,
                                          const unsigned (&w)[Row<D>::W]) {
    using R = Row<D>;
#pragma unroll
    for (int p = 0; p < R::N4; ++p) {
        st_stream16(tab + (size_t)p * 16u * (size_t)n_pad + (size_t)s * 16u,
                    make_uint4(w[4 * p], w[4 * p + 1], w[4 * p + 2], w[4 * p + 3]));
    }
    if constexpr (R::N2 != 0) {
        st_stream8(tab + (size_t)R::N4 * 16u * (size_t)n_pad + (size_t)s * 8u,
                   make_uint2(w[4 * R::N4], w[4 * R::N4 + 1]));
    }
    if constexpr (R::N1 != 0) {
        st_stream4(tab + ((size_t)R::N4 * 16u + (size_t)R::N2 * 8u) * (size_t)n_pad + (size_t)s * 4u, w[R::W - 1]);
    }
}

// ---------------------------------------------------------------------------
// Corner enumeration of the reference.
//   D == 2 : idxs[0..3] = (i0,i1) (i0,i1+1) (i0+1,i1) (i0+1,i1+1)   (:201-209)
//            -> bit of dim d in corner c is (c >> (1-d)) & 1
//   D >= 3 : bit d of c selects i[d] or i[d]+1                       (:602-613, :1030-1041)
// ---------------------------------------------------------------------------
template <int D>
__host__ __device__ constexpr int corner_bit(int c, int d) {
    return D == 2 ? ((c >> (1 - d)) & 1) : ((c >> d) & 1);
}

// Expected next-state value  sum_c w_c * V[idx_c]  with the reference's exact
// arithmetic: w_c = ((((1*f_0)*f_1)...)*f_{D-1}), f_d = bit ? frac_d : 1-frac_d,
// accumulated as ev = fmaf(w_c, V[idx_c], ev) for c = 0 .. 2^D-1 from ev = 0
// (:236-239, :643-646, :1074-1076).  Prefix products are shared between corners
// (same multiplication order per corner, so the same rounding).
template <int D>
__device__ __forceinline__ float expected_value(const float* __restrict__ V, int base,
                                                const float (&frac)[D], const int (&stride)[D]) {
    constexpr int C = 1 << D;
    if constexpr (D == 1) {
        const float* v = V + base;
        float ev = fmaf(1.0f - frac[0], v[0], 0.0f);
        return fmaf(frac[0], v[stride[0]], ev);
    } else if constexpr (D == 2) {
        const float g0 = 1.0f - frac[0], g1 = 1.0f - frac[1];
        const float* v = V + base;
        const float v00 = v[0], v01 = v[stride[1]], v10 = v[stride[0]], v11 = v[stride[0] + stride[1]];
        float ev = 0.0f;
        ev = fmaf(g0 * g1, v00, ev);
        ev = fmaf(g0 * frac[1], v01, ev);
        ev = fmaf(frac[0] * g1, v10, ev);
        ev = fmaf(frac[0] * frac[1], v11, ev);
        return ev;
    } else {
        // weights over the first D-1 dims (2^(D-1) prefix products), last dim applied on the fly
        constexpr int H = C / 2;
        float pre[H];
        pre[0] = 1.0f - frac[0];
        pre[1] = frac[0];
#pragma unroll
        for (int d = 1; d < D - 1; ++d) {
            const float g = 1.0f - frac[d];
#pragma unroll
            for (int c = (1 << d) - 1; c >= 0; --c) {
                const float p = pre[c];
                pre[c + (1 << d)] = p * frac[d];
                pre[c] = p * g;
            }
        }
        const float gl = 1.0f - frac[D - 1];
        const float* v = V + base;
        float val[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            int off = 0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (corner_bit<D>(c, d)) off += stride[d];
            val[c] = v[off];
        }
        float ev = 0.0f;
#pragma unroll
        for (int c = 0; c < H; ++c) ev = fmaf(pre[c] * gl, val[c], ev);
#pragma unroll
        for (int c = 0; c < H; ++c) ev = fmaf(pre[c] * frac[D - 1], val[c + H], ev);
        return ev;
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------
// Evaluation sweep: one Jacobi Bellman backup of every local state with the
// residual fused in (replaces policy_eval_kernel* + the max_abs_diff
// ReductionKernel, src/cuda_policy_iteration.py:212-242, :616-649, :1044-1079,
// :164-172).  Launched back to back; each launch reads the sweep counter from
// the control block, so a CUDA graph of identical nodes runs a whole sync
// interval.
// ---------------------------------------------------------------------------
struct EvalParams {
    const unsigned char* rows;  // compacted rows of the current policy (n_pad states)
    float* V0;                  // ping
    float* V1;                  // pong  (both full length n_states)
    Ctl* ctl;
    float* partial;             // per-block residual maxima (written on check sweeps only)
    long long n_local;
    long long n_pad;
    long long s_begin;  // first global state of this rank
    float gamma;
    int j;      // position of this launch inside its batch: sweep index = ctl->base + j
    int check;  // 1 if the host will examine the residual of this sweep (sync point, :325)
    int stride[kMaxDims];
};

template <int D>
__global__ void __launch_bounds__(kBlock) eval_sweep_kernel(const EvalParams p) {
    const Ctl* __restrict__ ctl = p.ctl;
    if (ctl->done) return;
    const int par = (ctl->base + p.j + ctl->parity0) & 1;
    const float* __restrict__ Vin = par ? p.V1 : p.V0;
    float* __restrict__ Vout = par ? p.V0 : p.V1;

    const long long s = (long long)blockIdx.x * kBlock + threadIdx.x;
    float res = 0.0f;
    if (s < p.n_local) {
        unsigned w[Row<D>::W];
        load_row<D>(p.rows, p.n_pad, s, w);
        const int base = (int)w[0];
        const float vold = Vin[p.s_begin + s];
        float vnew;
        if (base == PI_ROW_ABSORBING) {
            vnew = vold;  // :221 / :625 / :1053
        } else {
            float ev = 0.0f;
            if (base >= 0) {
                float frac[D];
                int stride[D];
#pragma unroll
                for (int d = 0; d < D; ++d) { frac[d] = __uint_as_float(w[1 + d]); stride[d] = p.stride[d]; }
                ev = expected_value<D>(Vin, base, frac, stride);
            }
            vnew = fmaf(p.gamma, ev, __uint_as_float(w[D + 1]));  // reward + gamma*ev contracts to one FMA
        }
        Vout[p.s_begin + s] = vnew;
        res = fabsf(vnew - vold);
    }
    if (!p.check) return;  // the reference reads the residual only at sync points (:325-326)

    // block residual -> partial[blockIdx]; reduced by eval_reduce_kernel (no atomics on the sweep path)
    __shared__ float s_red[kBlock / 32];
    res = warp_max(res);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = res;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < kBlock / 32 ? s_red[threadIdx.x] : 0.0f;
        r = warp_max(r);
        if (threadIdx.x == 0) p.partial[blockIdx.x] = r;
    }
}

// End of a batch of k sweeps: advance the sweep counter; if the last sweep was a
// check sweep, reduce the per-block residuals; if `decide`, apply the sync-point
// test of policy_evaluation (:325-331).  When sharded the residual is all-reduced
// between this kernel (decide = 0) and eval_decide_kernel.
__global__ void __launch_bounds__(1024) eval_reduce_kernel(Ctl* ctl, const float* partial, int n_partial, int k,
                                                          int has_check, int decide, float theta) {
    if (ctl->done) return;
    __shared__ float s_red[32];
    float r = 0.0f;
    if (has_check) {
        for (int i = threadIdx.x; i < n_partial; i += blockDim.x) r = fmaxf(r, partial[i]);
        r = warp_max(r);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = r;
        __syncthreads();
        if (threadIdx.x < 32) {
            r = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.0f;
            r = warp_max(r);
        }
    }
    if (threadIdx.x == 0) {
        ctl->base += k;
        if (has_check) {
            ctl->last_delta = r;
            if (decide) {
                ctl->check_delta = r;
                if (r < theta) { ctl->done = 1; ctl->conv_sweep = ctl->base - 1; }
            }
        }
    }
}

__global__ void eval_decide_kernel(Ctl* ctl, const float* delta_src, float theta) {
    if (ctl->done) return;
    const float d = *delta_src;
    ctl->check_delta = d;
    if (d < theta) {
        ctl->done = 1;
        ctl->conv_sweep = ctl->base - 1;
    }
}

__global__ void eval_begin_kernel(Ctl* ctl, int parity0) {
    ctl->base = 0;
    ctl->parity0 = parity0;
    ctl->done = 0;
    ctl->conv_sweep = -1;
    ctl->last_delta = __int_as_float(0x7f800000);
    ctl->check_delta = __int_as_float(0x7f800000);
}

// ---------------------------------------------------------------------------
// Improvement: per state, Q(s,a) for a = 0..A-1 in order, strict '>' from
// -1e30f with best_a = 0 (lowest index wins ties, NaN never wins), absorbing
// states untouched (replaces policy_improve_kernel*, :244-283, :651-691,
// :1081-1123) — fused with (i) the "policy changed" count that replaces
// old_policy.copy() + cp.all(==) (:340,:354) and (ii) compaction of the
// winning row into the evaluation table.
// ---------------------------------------------------------------------------
struct ImproveParams {
    const unsigned char* table;  // [A] row tables, action stride = W*4*n_pad bytes
    unsigned char* rows;         // out: compacted rows of the new policy
    const float* V;              // current value function (full length)
    int* policy;                 // local policy (n_local)
    unsigned int* partial;       // per-block count of states whose action changed
    long long n_local;
    long long n_pad;
    int n_actions;
    float gamma;
    int stride[kMaxDims];
};

template <int D>
__global__ void __launch_bounds__(kBlock) improve_kernel(const ImproveParams p) {
    const long long s = (long long)blockIdx.x * kBlock + threadIdx.x;
    int changed = 0;
    if (s < p.n_local) {
        constexpr int W = Row<D>::W;
        const size_t a_stride = (size_t)W * 4u * (size_t)p.n_pad;
        unsigned best_row[W];
        load_row<D>(p.table, p.n_pad, s, best_row);
        const int old_a = p.policy[s];
        if ((int)best_row[0] != PI_ROW_ABSORBING) {
            int stride[D];
#pragma unroll
            for (int d = 0; d < D; ++d) stride[d] = p.stride[d];
            float max_q = -1.0e30f;
            int best_a = 0;
            unsigned w[W];
#pragma unroll
            for (int k = 0; k < W; ++k) w[k] = best_row[k];
            for (int a = 0; a < p.n_actions; ++a) {
                if (a > 0) load_row<D>(p.table + (size_t)a * a_stride, p.n_pad, s, w);
                const int base = (int)w[0];
                float ev = 0.0f;
                if (base >= 0) {
                    float frac[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) frac[d] = __uint_as_float(w[1 + d]);
                    ev = expected_value<D>(p.V, base, frac, stride);
                }
                const float q = fmaf(p.gamma, ev, __uint_as_float(w[D + 1]));
                if (q > max_q) {
                    max_q = q;
                    best_a = a;
#pragma unroll
                    for (int k = 0; k < W; ++k) best_row[k] = w[k];
                }
            }
            p.policy[s] = best_a;
            changed = (best_a != old_a);
        }
        store_row<D>(p.rows, p.n_pad, s, best_row);
    }
    const unsigned m = __ballot_sync(0xffffffffu, changed);
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) p.partial[blockIdx.x] = (unsigned)s_cnt;
}

__global__ void __launch_bounds__(1024) count_reduce_kernel(Ctl* ctl, const unsigned int* partial, int n_partial) {
    __shared__ unsigned long long s_red[32];
    unsigned long long r = 0;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) r += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
        ctl->changed = t;
    }
}

// rows[s] = table[policy[s]][s]  (after pi_upload_policy and at start-up)
template <int D>
__global__ void __launch_bounds__(kBlock) compact_rows_kernel(const unsigned char* table, unsigned char* rows,
                                                              const int* policy, long long n_local,
                                                              long long n_pad) {
    const long long s = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (s >= n_local) return;
    constexpr int W = Row<D>::W;
    unsigned w[W];
    load_row<D>(table + (size_t)policy[s] * ((size_t)W * 4u * (size_t)n_pad), n_pad, s, w);
    store_row<D>(rows, n_pad, s, w);
}

// Expand compact rows to the reference's corner form (parity checks only).
template <int D>
__global__ void expand_rows_kernel(const unsigned char* table, long long n_pad, long long s0, long long count,
                                   GridDesc g, int* idx, float* wgt, float* reward, unsigned char* term) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    constexpr int W = Row<D>::W;
    constexpr int C = 1 << D;
    unsigned w[W];
    load_row<D>(table, n_pad, s0 + k, w);
    const int base = (int)w[0];
    if (reward) reward[k] = __uint_as_float(w[D + 1]);
    if (term) term[k] = base == PI_ROW_TERMINATED ? 1 : (base == PI_ROW_ABSORBING ? 2 : 0);
    for (int c = 0; c < C; ++c) {
        int off = 0;
        float wc = 1.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int bit = corner_bit<D>(c, d);
            const float f = __uint_as_float(w[1 + d]);
            off += bit * g.stride[d];
            wc *= bit ? f : (1.0f - f);
        }
        if (idx) idx[k * C + c] = base >= 0 ? base + off : base;
        if (wgt) wgt[k * C + c] = base >= 0 ? wc : 0.0f;
    }
}

__global__ void fill_masked_kernel(float* V0, float* V1, const unsigned char* mask, long long n, float value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && mask[i]) { V0[i] = value; V1[i] = value; }
}

}  // namespace pi
