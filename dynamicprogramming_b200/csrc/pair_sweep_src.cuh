// pair_sweep_src.cuh — CUDA source of the JIT generic evaluation sweeps: the gather sweep
// (policy_eval_kernel_4d/_6d of the reference, src/cuda_policy_iteration.py:616-649, :1044-1079,
// + the max|x-y| reduction :563-571, :987-995) compiled at run time by NVRTC for sm_100a with the
// grid strides baked in as immediates.  Two forms, one entry point (gp_sweep):
//
//   GP_SINGLE = 1  ONE state per thread, scalar math, the 2^D gathers issued in explicit groups of
//                  GP_G loads with the next group in flight while the fma chain consumes the current
//                  one — the default generic sweep of large 6-D grids (K5: 1.47 vs 1.68 ms per sweep
//                  for the ahead-of-time kernel; L1 data pipe 81 % busy)
//   GP_SINGLE = 0  TWO states per thread whose weight trees and fma chains run as one packed
//                  mul.rn.f32x2 / fma.rn.f32x2 stream (opt-in; the FMA pipe is not the limit)
//
// build.py turns this file into the string pi::kPairSweepSrc; the host prepends
//
//   GP_D        grid dimensions (>= 3)
//   GP_THREADS  threads per CTA (multiple of 32)      GP_MINB   CTAs per SM
//   GP_LV       trailing dimensions whose factors are applied per corner instead of being stored in the
//               weight tree (1..3): the tree holds 2^(D-LV) packed nodes, every corner pays LV multiplies.
//               Fewer tree registers leave more registers for gathers in flight (the sweep is latency-bound).
//   GP_G        gathers per explicitly scheduled group (0: leave the schedule to ptxas)
//   gp_off[]    V offset of corner c: sum of the storage strides of the set bits (bit d <-> dim d)
//
// Why the packed form.  The scalar sweep (pi::eval_sweep_kernel) spends 2^D scalar FMUL for the
// leaf weights, 2^D-4 for the tree, 2^D FFMA for the chain and ~2^D IMAD for the gather
// addresses per state — all on the FMA pipe, where a 3-register FFMA/FMUL/IMAD issues at half
// rate on sm_100.  Here the two states of a thread share every FMA-pipe instruction (each
// half of an f32x2 op rounds exactly like the scalar op, in the reference's order
// w_c = ((((f_0 f_1) f_2) ..) f_{D-1}), ev = fma(w_c, V_c, ev), c ascending :602-613, :643-646,
// :1030-1041, :1074-1076 => bit-identical V) and the gather addresses are base + immediate.
// The gathers themselves stay scalar and fully general: nothing is assumed about the policy
// or about which dimension is stored fastest.  Lane l of warp w owns states 64 w' + l and
// 64 w' + 32 + l, so every gather instruction is still issued by 32 consecutive states
// (same coalescing as the scalar sweep), rows are read from the engine's plain storage order.

typedef unsigned long long gp_u64;

struct GpCtl {   // == pi::Ctl
    int base, parity0, done, conv_sweep;
    float last_delta, check_delta;
    unsigned long long changed;
    unsigned int pad[8];
};


// == pi::PeerOut (pi_kernels.cuh): sharded runs store new values straight into the V buffers of the
// ranks that need them (CUDA IPC peer pointers over NVLink); n == 0 otherwise.
struct GpPeerOut {
    int n;
    int pad;
    float* V0[7];
    float* V1[7];
    long long lo[7];
    long long hi[7];
};
// unrolled over the 7 entries: a run-time-indexed loop over a kernel-parameter struct is compiled into local-memory copies
__device__ __forceinline__ void gp_store_peers(const GpPeerOut& po, bool out_is_V0, long long g, float v) {
#pragma unroll
    for (int r = 0; r < 7; ++r)
        if (r < po.n && g >= po.lo[r] && g < po.hi[r]) (out_is_V0 ? po.V0[r] : po.V1[r])[g] = v;
}

struct GpParams {
    const unsigned char* rows;   // compacted rows of the policy, 16/8/4-byte planes (pi::Row<D>)
    float* V0;
    float* V1;
    const GpCtl* ctl;
    float* partial;              // [gridDim.x]
    long long n_local;
    long long n_pad;
    long long s_begin;
    float gamma;
    int j;
    int check;
    GpPeerOut peers;
    int rot;                     // sharded runs: block visited first (blocks whose values peers need go first); 0 on one GPU
    int n_peer_blocks;           // ... and how many blocks from there on can hold states a peer needs (the rest is interior)
    int sched_begin;             // one state per thread: this launch sweeps blocks [sched_begin, sched_begin + gridDim.x) of the (rotated)
    int n_blocks;                // schedule of n_blocks blocks (sharded runs with the DMA exchange launch boundary and interior separately)
    float2* P0;                  // GP_PAIRV: pair shadow of V0, P0[i] = (V0[i], V0[i+1]); else unused
    float2* P1;                  //           ... of V1
};

#define GP_W (GP_D + 2)
#define GP_N4 (GP_W / 4)
#define GP_N2 ((GP_W % 4) / 2)
#define GP_C (1 << GP_D)
#define GP_H (GP_C / 2)
#ifndef GP_LV
#define GP_LV 1
#endif
#define GP_NT (GP_C >> GP_LV)
#ifndef GP_G
#define GP_G 0        // gathers per explicitly scheduled group (0: leave the schedule to ptxas)
#endif

__device__ __forceinline__ gp_u64 gp_pk(float a, float b) {
    gp_u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void gp_unpk(gp_u64 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ gp_u64 gp_mul2(gp_u64 a, gp_u64 b) {
    gp_u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// same multiply, opaque to common-subexpression elimination (GP_LV >= 2: the partial products of the
// trailing dimensions are RE-computed per corner, not kept in registers)
__device__ __forceinline__ gp_u64 gp_mul2_nocse(gp_u64 a, gp_u64 b) {
    gp_u64 r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float gp_mul_nocse(float a, float b) {
    float r;
    asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float gp_ld_ordered(const float* p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ gp_u64 gp_fma2(gp_u64 a, gp_u64 b, gp_u64 c) {
    gp_u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float gp_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// row of local state s (streaming: read once, keep V in L1/L2)
__device__ __forceinline__ void gp_load_row(const unsigned char* __restrict__ tab, long long n_pad, long long s,
                                            unsigned (&w)[GP_W]) {
#pragma unroll
    for (int q = 0; q < GP_N4; ++q) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(tab + (size_t)q * 16u * (size_t)n_pad + (size_t)s * 16u));
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
#if GP_N2
    {
        uint2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y)
                     : "l"(tab + (size_t)GP_N4 * 16u * (size_t)n_pad + (size_t)s * 8u));
        w[4 * GP_N4] = v.x; w[4 * GP_N4 + 1] = v.y;
    }
#endif
#if GP_W % 2
    {
        unsigned v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v)
                     : "l"(tab + ((size_t)GP_N4 * 16u + (size_t)GP_N2 * 8u) * (size_t)n_pad + (size_t)s * 4u));
        w[GP_W - 1] = v;
    }
#endif
}

#ifndef GP_SINGLE
#define GP_SINGLE 0   // 1: one state per thread (scalar math), same immediates and explicit gather schedule
#endif

#ifndef GP_PAIRV
#define GP_PAIRV 0    // 1 (GP_SINGLE only, fast-stored dimension = logical dimension 0): gather from a PAIR SHADOW of V
#endif

#if GP_PAIRV
// P[i] = (V[i], V[i+1]): the two corners of a cell along the fast dimension are one aligned 64-bit load instead of
// two 32-bit loads that touch the same lines — half the gather requests.  Modelled on the converged K5 policy:
// 5.0 instead of 7.9 L1 wavefronts per window per warp.  The sweep reads pairs, and writes the plain
// value (everything else in the engine reads plain V) plus the two pair halves that contain it.
__device__ __forceinline__ float2 gp_ld_pair_ordered(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
#endif

#if GP_SINGLE
// One state per thread: the scalar sweep (pi::eval_sweep_kernel + expected_value_grouped) with the grid's
// strides as immediates — no address arithmetic, fewer registers, more resident warps.
extern "C" __global__ void __launch_bounds__(GP_THREADS, GP_MINB) gp_sweep(const GpParams p)
{
    const GpCtl* __restrict__ ctl = p.ctl;
    if (ctl->done) return;
    const int par = (ctl->base + p.j + ctl->parity0) & 1;
    const float* __restrict__ Vin = par ? p.V1 : p.V0;
    float* __restrict__ Vout = par ? p.V0 : p.V1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned sched = blockIdx.x + (unsigned)p.sched_begin;
    const unsigned blk = sched + (unsigned)p.rot < (unsigned)p.n_blocks ? sched + (unsigned)p.rot : sched + (unsigned)p.rot - (unsigned)p.n_blocks;
    const long long s = (long long)blk * GP_THREADS + threadIdx.x;
    // peers need values only from the first blocks of the (rotated) schedule
    const bool peer_blk = sched < (unsigned)p.n_peer_blocks;
#if GP_PAIRV
    const float2* __restrict__ Pin = par ? p.P1 : p.P0;
    float2* __restrict__ Pout = par ? p.P0 : p.P1;
#endif
    float res = 0.0f;
    if (s < p.n_local) {
        unsigned w[GP_W];
        gp_load_row(p.rows, p.n_pad, s, w);
        const int b = (int)w[0];
        const float vold = Vin[p.s_begin + s];
        float vnew;
        if (b == -2) {
            vnew = vold;
        } else {
            float ev = 0.0f;
            if (b >= 0) {
                const float* v = Vin + b;
                float pre[GP_NT];
                pre[0] = 1.0f - __uint_as_float(w[1]);
                pre[1] = __uint_as_float(w[1]);
#pragma unroll
                for (int d = 1; d < GP_D - GP_LV; ++d) {
                    const float f = __uint_as_float(w[1 + d]), g = 1.0f - f;
#pragma unroll
                    for (int c = GP_NT / 2 - 1; c >= 0; --c) {
                        if (c < (1 << d)) {
                            const float t = pre[c];
                            pre[c + (1 << d)] = t * f;
                            pre[c] = t * g;
                        }
                    }
                }
                float wt[GP_LV][2];
#pragma unroll
                for (int k = 0; k < GP_LV; ++k) {
                    const float l = __uint_as_float(w[1 + GP_D - GP_LV + k]);
                    wt[k][0] = 1.0f - l;
                    wt[k][1] = l;
                }
#if GP_G == 0
#pragma unroll
                for (int c = 0; c < GP_C; ++c) {
                    float leaf = pre[c & (GP_NT - 1)];
#pragma unroll
                    for (int k = 0; k < GP_LV; ++k) leaf = leaf * wt[k][(c >> (GP_D - GP_LV + k)) & 1];
                    ev = fmaf(leaf, __ldg(v + gp_off[c]), ev);
                }
#else
                float buf[2][GP_G];
#if GP_PAIRV
                // corners 2o and 2o+1 differ in bit 0 <-> the fast dimension (stride 1): one pair load serves both
                const float2* pv = Pin + b;
#pragma unroll
                for (int i = 0; i < GP_G; i += 2) {
                    const float2 t = gp_ld_pair_ordered(pv + gp_off[i]);
                    buf[0][i] = t.x; buf[0][i + 1] = t.y;
                }
#else
#pragma unroll
                for (int i = 0; i < GP_G; ++i) buf[0][i] = gp_ld_ordered(v + gp_off[i]);
#endif
#pragma unroll
                for (int g = 0; g < GP_C / GP_G; ++g) {
                    if (g + 1 < GP_C / GP_G) {
#if GP_PAIRV
#pragma unroll
                        for (int i = 0; i < GP_G; i += 2) {
                            const float2 t = gp_ld_pair_ordered(pv + gp_off[(g + 1) * GP_G + i]);
                            buf[(g + 1) & 1][i] = t.x; buf[(g + 1) & 1][i + 1] = t.y;
                        }
#else
#pragma unroll
                        for (int i = 0; i < GP_G; ++i) buf[(g + 1) & 1][i] = gp_ld_ordered(v + gp_off[(g + 1) * GP_G + i]);
#endif
                    }
#pragma unroll
                    for (int i = 0; i < GP_G; ++i) {
                        const int c = g * GP_G + i;
                        float leaf = pre[c & (GP_NT - 1)];
#pragma unroll
                        for (int k = 0; k < GP_LV; ++k) {
                            const float f = wt[k][(c >> (GP_D - GP_LV + k)) & 1];
                            leaf = (k + 1 < GP_LV) ? gp_mul_nocse(leaf, f) : leaf * f;
                        }
                        ev = fmaf(leaf, buf[g & 1][i], ev);
                    }
                }
#endif
            }
            vnew = fmaf(p.gamma, ev, __uint_as_float(w[GP_D + 1]));
        }
        Vout[p.s_begin + s] = vnew;
#if GP_PAIRV
        {   // the two pairs that contain this value: (V[g], V[g+1]).x and (V[g-1], V[g]).y
            const long long g = p.s_begin + s;
            Pout[g].x = vnew;
            if (g > 0) Pout[g - 1].y = vnew;
        }
#endif
        if (peer_blk) gp_store_peers(p.peers, par != 0, p.s_begin + s, vnew);
        res = fabsf(vnew - vold);
    }
    if (!p.check) return;
    __shared__ float s_red[32];
    res = gp_warp_max(res);
    if (lane == 0) s_red[warp] = res;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < GP_THREADS / 32 ? s_red[threadIdx.x] : 0.0f;
        r = gp_warp_max(r);
        if (threadIdx.x == 0) p.partial[sched] = r;
    }
}
#else
extern "C" __global__ void __launch_bounds__(GP_THREADS, GP_MINB) gp_sweep(const GpParams p)
{
    const GpCtl* __restrict__ ctl = p.ctl;
    if (ctl->done) return;
    const int par = (ctl->base + p.j + ctl->parity0) & 1;
    const float* __restrict__ Vin = par ? p.V1 : p.V0;
    float* __restrict__ Vout = par ? p.V0 : p.V1;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long sA = ((long long)blockIdx.x * (GP_THREADS / 32) + warp) * 64 + lane;
    const long long sB = sA + 32;
    const bool inA = sA < p.n_local, inB = sB < p.n_local;
    float res = 0.0f;

    if (inA) {
        unsigned wA[GP_W], wB[GP_W];
        gp_load_row(p.rows, p.n_pad, sA, wA);
        if (inB) gp_load_row(p.rows, p.n_pad, sB, wB);
        else {
#pragma unroll
            for (int k = 0; k < GP_W; ++k) wB[k] = k == 0 ? 0xfffffffeu : 0u;
        }
        const float voldA = Vin[p.s_begin + sA];
        const float voldB = inB ? Vin[p.s_begin + sB] : 0.0f;
        const int bA = (int)wA[0], bB = (int)wB[0];
        // sentinel rows gather from a valid address and are discarded afterwards
        const float* vA = Vin + (bA >= 0 ? bA : 0);
        const float* vB = Vin + (bB >= 0 ? bB : 0);

        // weight tree over dims 0..D-1-LV, packed (state A, state B)
        gp_u64 node[GP_NT];
        {
            const float fa = __uint_as_float(wA[1]), fb = __uint_as_float(wB[1]);
            node[0] = gp_pk(1.0f - fa, 1.0f - fb);
            node[1] = gp_pk(fa, fb);
#pragma unroll
            for (int d = 1; d < GP_D - GP_LV; ++d) {
                const float xa = __uint_as_float(wA[1 + d]), xb = __uint_as_float(wB[1 + d]);
                const gp_u64 f = gp_pk(xa, xb), g = gp_pk(1.0f - xa, 1.0f - xb);
#pragma unroll
                for (int c = GP_NT / 2 - 1; c >= 0; --c) {   // constant trip count: node[] stays in registers
                    if (c < (1 << d)) {
                        const gp_u64 t = node[c];
                        node[c + (1 << d)] = gp_mul2(t, f);
                        node[c] = gp_mul2(t, g);
                    }
                }
            }
        }
        // factors of the trailing dimensions D-LV .. D-1: wt[k][bit]
        gp_u64 wt[GP_LV][2];
#pragma unroll
        for (int k = 0; k < GP_LV; ++k) {
            const float la = __uint_as_float(wA[1 + GP_D - GP_LV + k]), lb = __uint_as_float(wB[1 + GP_D - GP_LV + k]);
            wt[k][0] = gp_pk(1.0f - la, 1.0f - lb);
            wt[k][1] = gp_pk(la, lb);
        }

        gp_u64 ev = gp_pk(0.0f, 0.0f);
#if GP_G == 0
#pragma unroll
        for (int c = 0; c < GP_C; ++c) {
            gp_u64 leaf = node[c & (GP_NT - 1)];
#pragma unroll
            for (int k = 0; k < GP_LV; ++k) {
                const gp_u64 f = wt[k][(c >> (GP_D - GP_LV + k)) & 1];
                leaf = (k + 1 < GP_LV) ? gp_mul2_nocse(leaf, f) : gp_mul2(leaf, f);
            }
            ev = gp_fma2(leaf, gp_pk(__ldg(vA + gp_off[c]), __ldg(vB + gp_off[c])), ev);
        }
#else
        // explicitly scheduled gathers (see pi::expected_value_grouped): groups of GP_G corners, the loads of
        // group g+1 are issued (volatile: in order) before the fma chain of group g consumes its values
        float bufA[2][GP_G], bufB[2][GP_G];
#pragma unroll
        for (int i = 0; i < GP_G; ++i) { bufA[0][i] = gp_ld_ordered(vA + gp_off[i]); bufB[0][i] = gp_ld_ordered(vB + gp_off[i]); }
#pragma unroll
        for (int g = 0; g < GP_C / GP_G; ++g) {
            if (g + 1 < GP_C / GP_G) {
#pragma unroll
                for (int i = 0; i < GP_G; ++i) {
                    bufA[(g + 1) & 1][i] = gp_ld_ordered(vA + gp_off[(g + 1) * GP_G + i]);
                    bufB[(g + 1) & 1][i] = gp_ld_ordered(vB + gp_off[(g + 1) * GP_G + i]);
                }
            }
#pragma unroll
            for (int i = 0; i < GP_G; ++i) {
                const int c = g * GP_G + i;
                gp_u64 leaf = node[c & (GP_NT - 1)];
#pragma unroll
                for (int k = 0; k < GP_LV; ++k) {
                    const gp_u64 f = wt[k][(c >> (GP_D - GP_LV + k)) & 1];
                    leaf = (k + 1 < GP_LV) ? gp_mul2_nocse(leaf, f) : gp_mul2(leaf, f);
                }
                ev = gp_fma2(leaf, gp_pk(bufA[g & 1][i], bufB[g & 1][i]), ev);
            }
        }
#endif
        float evA, evB;
        gp_unpk(ev, evA, evB);
        // sentinel rows: terminated (-1) -> sum := 0 (:231-232); absorbing (-2) -> new_V := V (:221)
        const float vnewA = bA == -2 ? voldA : fmaf(p.gamma, bA >= 0 ? evA : 0.0f, __uint_as_float(wA[GP_D + 1]));
        Vout[p.s_begin + sA] = vnewA;
        if (p.peers.n) gp_store_peers(p.peers, par != 0, p.s_begin + sA, vnewA);
        res = fabsf(vnewA - voldA);
        if (inB) {
            const float vnewB = bB == -2 ? voldB : fmaf(p.gamma, bB >= 0 ? evB : 0.0f, __uint_as_float(wB[GP_D + 1]));
            Vout[p.s_begin + sB] = vnewB;
            if (p.peers.n) gp_store_peers(p.peers, par != 0, p.s_begin + sB, vnewB);
            res = fmaxf(res, fabsf(vnewB - voldB));
        }
    }
    if (!p.check) return;
    __shared__ float s_red[32];
    res = gp_warp_max(res);
    if (lane == 0) s_red[warp] = res;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < GP_THREADS / 32 ? s_red[threadIdx.x] : 0.0f;
        r = gp_warp_max(r);
        if (threadIdx.x == 0) p.partial[blockIdx.x] = r;
    }
}
#endif   // GP_SINGLE
