/*
 * oracle/pi_oracle.c — CPU restatement of the reference's policy-iteration
 * algorithm (nicoRomeroCuruchet/DynamicProgramming, src/cuda_policy_iteration.py).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker / baseline.
 * The product (dynamicprogramming_b200/) never links or calls this file.
 *
 * Parity status: PINNED AT TOLERANCE against the reference's two golden
 * artefacts (runners/results/{mountain_car,continuous_mountain_car}_cuda_policy.npz,
 * reduced copies under tests/golden/): float32 CPU arithmetic cannot be
 * bit-identical to the GPU (libm vs CUDA sinf/cosf, compiler-chosen FMA
 * contraction inside step_dynamics).  The bit-exact oracle is the reference's
 * own kernels compiled into oracle/_ref/ and run on the GPU (oracle/ref_runner.py).
 *
 * Like the reference, and unlike the product, nothing is tabulated: every
 * sweep re-runs step_dynamics and the multilinear lookup for every state.
 * The environment's step function is passed in as a pointer (built from the
 * plugin's `step_dynamics` source by oracle/build_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_DIMS 6
#define ORACLE_MAX_CORNERS 64

typedef void (*oracle_step_fn)(const float* s, float action, float* ns, float* reward, int* terminated);

typedef struct {
    int n_dims;
    int shape[ORACLE_MAX_DIMS];
    int strides[ORACLE_MAX_DIMS];
    float lo[ORACLE_MAX_DIMS];
    float hi[ORACLE_MAX_DIMS];
} oracle_grid;

/* get_barycentric_2d (src/cuda_policy_iteration.py:183-210),
 * get_barycentric_4d (:580-614), get_barycentric_6d (:1007-1042).
 * n = (s-lo)/(hi-lo)*(shape-1); clamp to [0, shape-1]; i = min((int)n, shape-2);
 * frac = n - i.  2-D corner order: (i0,i1) (i0,i1+1) (i0+1,i1) (i0+1,i1+1);
 * N-D: bit d of corner c selects i[d] or i[d]+1, weight = ((1*f0)*f1)*... */
void oracle_barycentric(const oracle_grid* g, const float* s, int32_t* idxs, float* wgts) {
    const int D = g->n_dims;
    int i[ORACLE_MAX_DIMS];
    float frac[ORACLE_MAX_DIMS];
    for (int d = 0; d < D; ++d) {
        float n = (s[d] - g->lo[d]) / (g->hi[d] - g->lo[d]) * (float)(g->shape[d] - 1);
        n = fmaxf(0.0f, fminf(n, (float)(g->shape[d] - 1)));
        int id = (int)n;
        if (id > g->shape[d] - 2) id = g->shape[d] - 2;
        i[d] = id;
        frac[d] = n - (float)id;
    }
    if (D == 2) {
        const float d0 = frac[0], d1 = frac[1];
        idxs[0] = i[0] * g->strides[0] + i[1] * g->strides[1];
        idxs[1] = i[0] * g->strides[0] + (i[1] + 1) * g->strides[1];
        idxs[2] = (i[0] + 1) * g->strides[0] + i[1] * g->strides[1];
        idxs[3] = (i[0] + 1) * g->strides[0] + (i[1] + 1) * g->strides[1];
        wgts[0] = (1.0f - d0) * (1.0f - d1);
        wgts[1] = (1.0f - d0) * d1;
        wgts[2] = d0 * (1.0f - d1);
        wgts[3] = d0 * d1;
        return;
    }
    const int C = 1 << D;
    for (int c = 0; c < C; ++c) {
        int idx = 0;
        float wgt = 1.0f;
        for (int d = 0; d < D; ++d) {
            const int bit = (c >> d) & 1;
            idx += (i[d] + bit) * g->strides[d];
            wgt *= bit ? frac[d] : (1.0f - frac[d]);
        }
        idxs[c] = idx;
        wgts[c] = wgt;
    }
}

/* Q(s, a) = reward + gamma * sum_c w_c V[idx_c]; the sum is an fmaf chain from
 * 0 in ascending corner order and the last step is one fused multiply-add, as
 * NVRTC compiles :236-241, :643-648, :1074-1078. */
static float oracle_q(const oracle_grid* g, oracle_step_fn step, const float* s, float action, const float* V,
                      float gamma) {
    float ns[ORACLE_MAX_DIMS], reward;
    int terminated = 0;
    step(s, action, ns, &reward, &terminated);
    float ev = 0.0f;
    if (!terminated) {
        int32_t idxs[ORACLE_MAX_CORNERS];
        float wgts[ORACLE_MAX_CORNERS];
        oracle_barycentric(g, ns, idxs, wgts);
        const int C = 1 << g->n_dims;
        for (int c = 0; c < C; ++c) ev = fmaf(wgts[c], V[idxs[c]], ev);
    }
    return fmaf(gamma, ev, reward);
}

/* Transition rows for one action (what the product tabulates): idx[n][C], w[n][C],
 * reward[n], terminated[n] — step_dynamics + get_barycentric for every state. */
void oracle_rows(const oracle_grid* g, oracle_step_fn step, const float* states, int64_t n_states, float action,
                 int32_t* idx, float* w, float* reward, uint8_t* terminated, float* next_states) {
    const int D = g->n_dims, C = 1 << D;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n_states; ++s) {
        float ns[ORACLE_MAX_DIMS], r;
        int t = 0;
        step(states + s * D, action, ns, &r, &t);
        oracle_barycentric(g, ns, idx + s * C, w + s * C);
        reward[s] = r;
        terminated[s] = (uint8_t)(t != 0);
        if (next_states) memcpy(next_states + s * D, ns, sizeof(float) * D);
    }
}

/* policy_eval_kernel{,_4d,_6d} (:212-242, :616-649, :1044-1079) followed by the
 * max|new_V - V| reduction (:164-172).  Returns the residual. */
float oracle_eval_sweep(const oracle_grid* g, oracle_step_fn step, const float* states, const float* actions,
                        const int32_t* policy, const float* V, float* new_V, const uint8_t* is_term,
                        int64_t n_states, float gamma) {
    const int D = g->n_dims;
    float delta = 0.0f;
#pragma omp parallel for schedule(static) reduction(max : delta)
    for (int64_t s = 0; s < n_states; ++s) {
        float v;
        if (is_term[s]) {
            v = V[s];
        } else {
            v = oracle_q(g, step, states + s * D, actions[policy[s]], V, gamma);
        }
        new_V[s] = v;
        const float d = fabsf(v - V[s]);
        if (d > delta) delta = d;
    }
    return delta;
}

/* policy_improve_kernel{,_4d,_6d} (:244-283, :651-691, :1081-1123): strict '>'
 * from -1e30f, best_a = 0 -> lowest index wins ties; terminal states untouched.
 * Returns the number of states whose action changed (stable <=> 0, :340,:354). */
int64_t oracle_improve(const oracle_grid* g, oracle_step_fn step, const float* states, const float* actions,
                       int n_actions, int32_t* policy, const float* V, const uint8_t* is_term, int64_t n_states,
                       float gamma) {
    const int D = g->n_dims;
    int64_t changed = 0;
#pragma omp parallel for schedule(static) reduction(+ : changed)
    for (int64_t s = 0; s < n_states; ++s) {
        if (is_term[s]) continue;
        float max_q = -1.0e30f;
        int best_a = 0;
        for (int a = 0; a < n_actions; ++a) {
            const float q = oracle_q(g, step, states + s * D, actions[a], V, gamma);
            if (q > max_q) { max_q = q; best_a = a; }
        }
        if (policy[s] != best_a) ++changed;
        policy[s] = best_a;
    }
    return changed;
}

/* policy_evaluation (:300-336): residual examined only when i % 25 == 0 or on the
 * last iteration; returns at the first such i with delta < theta.  V / new_V are
 * swapped in place (the caller's V always holds the newest values on return).
 * Writes the number of sweeps executed to *sweeps; returns the last residual read. */
float oracle_policy_evaluation(const oracle_grid* g, oracle_step_fn step, const float* states, const float* actions,
                               const int32_t* policy, float* V, float* scratch, const uint8_t* is_term,
                               int64_t n_states, float gamma, float theta, int max_eval_iter, int sync_interval,
                               int* sweeps) {
    float delta = INFINITY;
    float* cur = V;
    float* nxt = scratch;
    int i = 0;
    int done = 0;
    for (i = 0; i < max_eval_iter; ++i) {
        const float d = oracle_eval_sweep(g, step, states, actions, policy, cur, nxt, is_term, n_states, gamma);
        float* t = cur; cur = nxt; nxt = t;
        if (i % sync_interval == 0 || i == max_eval_iter - 1) {
            delta = d;
            if (delta < theta) { done = 1; break; }
        }
    }
    if (sweeps) *sweeps = done ? i + 1 : max_eval_iter;
    if (cur != V) memcpy(V, cur, sizeof(float) * (size_t)n_states);
    return delta;
}

/* run (:357-370): evaluation / improvement until the policy is stable or
 * max_pi_iter is hit; V is warm-started across iterations.  Returns the number
 * of PI iterations executed (negative if the policy never became stable). */
int oracle_run(const oracle_grid* g, oracle_step_fn step, const float* states, const float* actions, int n_actions,
               int32_t* policy, float* V, const uint8_t* is_term, int64_t n_states, float gamma, float theta,
               int max_eval_iter, int max_pi_iter, int sync_interval, int64_t* total_sweeps) {
    float* scratch = (float*)malloc(sizeof(float) * (size_t)n_states);
    memcpy(scratch, V, sizeof(float) * (size_t)n_states);
    int64_t tot = 0;
    int n, stable = 0;
    for (n = 0; n < max_pi_iter; ++n) {
        int sw = 0;
        oracle_policy_evaluation(g, step, states, actions, policy, V, scratch, is_term, n_states, gamma, theta,
                                 max_eval_iter, sync_interval, &sw);
        tot += sw;
        if (oracle_improve(g, step, states, actions, n_actions, policy, V, is_term, n_states, gamma) == 0) {
            stable = 1;
            ++n;
            break;
        }
    }
    free(scratch);
    if (total_sweeps) *total_sweeps = tot;
    return stable ? n : -n;
}

/* utils/barycentric.py:11-73 as numba actually types it (verified with
 * inspect_types and pinned by tests/golden/barycentric_inference_golden.npz):
 * step_sizes is a FLOAT32 array (the array expression is evaluated per element
 * as float32-float32 -> float32, / int64 -> float64, then stored as float32);
 * cell = (p-lo)/step is float32; t = (p - (lo + idx*step))/step is evaluated in
 * float64 (int64*float32 promotes) and stored as float32; weights accumulate in
 * float64 and are stored as float32; corner order follows corner_bits. */
__attribute__((optimize("fp-contract=off")))  /* numba/LLVM does not contract */
void oracle_inference_weights(const oracle_grid* g, const int32_t* corner_bits, const float* points,
                              int64_t n_points, float* weights, int32_t* indices) {
    const int D = g->n_dims, C = 1 << D;
    for (int64_t p = 0; p < n_points; ++p) {
        int base[ORACLE_MAX_DIMS];
        float t[ORACLE_MAX_DIMS];
        for (int d = 0; d < D; ++d) {
            const float step = (float)((double)(float)(g->hi[d] - g->lo[d]) / (double)(g->shape[d] - 1));
            float x = points[p * D + d];
            float q = x < g->hi[d] ? x : g->hi[d];
            q = q > g->lo[d] ? q : g->lo[d];
            const float cell = (q - g->lo[d]) / step;
            int id = (int)cell;
            if (id >= g->shape[d] - 1) id = g->shape[d] - 2;
            base[d] = id;
            t[d] = (float)(((double)q - ((double)g->lo[d] + (double)id * (double)step)) / (double)step);
        }
        for (int c = 0; c < C; ++c) {
            double w = 1.0;
            int flat = 0;
            for (int d = 0; d < D; ++d) {
                const int bit = corner_bits[c * D + d];
                w *= bit ? (double)t[d] : (1.0 - (double)t[d]);
                flat += (base[d] + bit) * g->strides[d];
            }
            weights[p * C + c] = (float)w;
            indices[p * C + c] = flat;
        }
    }
}
