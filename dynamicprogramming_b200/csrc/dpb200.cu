// dpb200.cu — host side of libdpb200.so: the C ABI declared in include/dpb200.h.
//
// Owns device memory, the NVRTC-compiled transition-table builder, the CUDA
// graphs that run a whole sync interval of evaluation sweeps without host
// involvement, and (when sharded) the NCCL communicator used to exchange V.
// The policy-iteration control flow mirrors src/cuda_policy_iteration.py:300-370
// of the reference exactly (sync every 25 sweeps, warm-started V, for/else
// warning semantics); only the arithmetic lives in pi_kernels.cuh.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dpb200.h"
#include "pi_kernels.cuh"
#include "table_build_src.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(PI_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                   \
    } while (0)

// ---------------------------------------------------------------------------
// NCCL, loaded lazily so the library loads on a box without NCCL/GPU.
// ---------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat32 = 7, ncclInt64 = 4, ncclUint64 = 5, ncclInt32 = 2, ncclUint8 = 1 };
enum { ncclSum = 0, ncclMax = 2 };
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return false;
#define L(sym) *(void**)(&sym) = dlsym(h, "nccl" #sym)
        L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(AllReduce); L(Send); L(Recv); L(GroupStart); L(GroupEnd);
        L(GetErrorString);
#undef L
        return CommInitRank && AllReduce && Send && Recv && GroupStart && GroupEnd;
    }
};
Nccl g_nccl;

#define NC(call)                                                                                  \
    do {                                                                                          \
        int _r = (call);                                                                          \
        if (_r != 0)                                                                              \
            return fail(PI_ERR_COMM, "%s failed: %s", #call,                                      \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");        \
    } while (0)

// Parameter block of the NVRTC kernel; must match PiBuildParams in table_build_src.h
// (the JIT source declares the arrays with PI_D entries, so pack accordingly).
struct BuildParamsHost {
    std::vector<unsigned char> bytes;
};

template <typename T>
void put(std::vector<unsigned char>& b, const T& v, size_t align) {
    while (b.size() % align) b.push_back(0);
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
    b.insert(b.end(), p, p + sizeof(T));
}

}  // namespace

struct pi_engine {
    int D = 0, A = 0;
    long long N = 0;
    pi::GridDesc g{};
    pi_config cfg{};
    int device = 0;
    int rank = 0, world = 1;
    long long s_begin = 0, s_end = 0, n_local = 0, n_pad = 0;

    cudaStream_t stream = nullptr;
    float* d_V[2] = {nullptr, nullptr};
    int cur = 0;  // index of the buffer that holds the current value function
    int* d_policy = nullptr;
    unsigned char* d_term = nullptr;       // local terminal mask
    unsigned char* d_mask_full = nullptr;  // scratch for set_terminal / set_values (N bytes)
    void* d_scratch = nullptr;             // N*4 bytes, reference-order staging when the layout is permuted
    int fast_dim = -1;                     // logical dimension stored with stride 1
    bool identity_layout = true;
    int sm_count = 148;
    int blocks_per_sm = 4;                 // resident evaluation blocks per SM (occupancy query)
    int lookahead = 0;                     // L2 row prefetch distance in blocks (DPB200_LOOKAHEAD, default = resident blocks)
    bool pair_kernel = false;              // evaluation sweeps use eval_sweep_pair_kernel<D, fast_dim>
    double probe_lines[PI_MAX_DIMS] = {};  // layout probe: avg 128-B lines per warp gather, per candidate
    unsigned char* d_table = nullptr;
    unsigned char* d_rows = nullptr;
    float* d_actions = nullptr;
    float* d_axes[PI_MAX_DIMS] = {};
    pi::Ctl* d_ctl = nullptr;
    pi::Ctl* h_ctl = nullptr;  // pinned, 4 slots
    float* d_partial = nullptr;  // per-block residual maxima / change counts
    float* d_delta_g = nullptr;  // global residual after all-reduce (sharded)
    unsigned long long* d_changed_g = nullptr;

    cudaLibrary_t lib = nullptr;
    cudaKernel_t build_kernel = nullptr;
    std::string nvrtc_log;

    // graphs: [start parity][0: one sweep + decide, 1: sync_interval sweeps + decide, 2: sync sweeps, no decide]
    cudaGraphExec_t graphs[2][3] = {};
    cudaEvent_t ev[4] = {};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;

    bool table_built = false;
    bool results_gathered = false;
    pi_stats stats{};
    long long launches = 0;

    pi_log_fn log_fn = nullptr;
    void* log_user = nullptr;

    // sharded exchange: for each peer, the sub-range of its slice this rank needs
    ncclComm_t comm = nullptr;
    std::vector<long long> need_lo, need_hi;   // what I need from peer r  (global indices)
    std::vector<long long> give_lo, give_hi;   // what peer r needs from me
    bool exchange_ready = false;
};

namespace {

void logf(pi_engine* e, int level, const char* fmt, ...) {
    if (!e->log_fn) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    e->log_fn(level, buf, e->log_user);
}

inline unsigned nblocks(long long n) { return (unsigned)((n + pi::kBlock - 1) / pi::kBlock); }

// ------------------------------------------------------------------ NVRTC --
int compile_to_cubin(int D, const char* dynamics_src, std::vector<char>& cubin, std::string& log) {
    std::string src;
    src += "#define PI_D " + std::to_string(D) + "\n";
    // call text: step_dynamics(x0.., action, &nx0.., &reward, &terminated)
    std::string call = "step_dynamics(";
    for (int d = 0; d < D; ++d) call += "pi_x[" + std::to_string(d) + "], ";
    call += "pi_action, ";
    for (int d = 0; d < D; ++d) call += "&pi_nx[" + std::to_string(d) + "], ";
    call += "&pi_reward, &pi_terminated)";
    src += "#define PI_CALL_STEP_DYNAMICS " + call + "\n";
    src += dynamics_src;
    src += "\n";
    src += pi::kTableBuildSrc;

    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, src.c_str(), "pi_table_build.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return fail(PI_ERR_COMPILE, "nvrtcCreateProgram: %s", nvrtcGetErrorString(r));
    // Same defaults as cupy.RawModule: architecture only (fmad=true, prec-div=true, no fast math).
    const char* opts[] = {"--gpu-architecture=sm_100a"};
    r = nvrtcCompileProgram(prog, 1, opts);
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    log.assign(log_size, '\0');
    if (log_size) nvrtcGetProgramLog(prog, &log[0]);
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(PI_ERR_COMPILE, "NVRTC failed to compile the dynamics source (%s):\n%s",
                    nvrtcGetErrorString(r), log.c_str());
    }
    size_t n = 0;
    nvrtcGetCUBINSize(prog, &n);
    cubin.resize(n);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return PI_OK;
}

int compile_builder(pi_engine* e, const char* dynamics_src) {
    std::vector<char> cubin;
    int rc = compile_to_cubin(e->D, dynamics_src, cubin, e->nvrtc_log);
    if (rc) return rc;
    CU(cudaLibraryLoadData(&e->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    CU(cudaLibraryGetKernel(&e->build_kernel, e->lib, "pi_build_rows"));
    return PI_OK;
}

// ------------------------------------------------------------ table build --
void set_layout(pi_engine* e, int fast_dim) {
    const int D = e->D;
    int k = 0;
    for (int d = 0; d < D; ++d)
        if (d != fast_dim) e->g.perm[k++] = d;
    e->g.perm[D - 1] = fast_dim;
    long long st = 1;
    for (int pos = D - 1; pos >= 0; --pos) {
        e->g.istride[e->g.perm[pos]] = (int)st;
        st *= e->g.shape[e->g.perm[pos]];
    }
    e->fast_dim = fast_dim;
    e->identity_layout = (fast_dim == D - 1);
}

// Launch the NVRTC-compiled builder for `n_states` internal states starting at `s_begin`.
int launch_build(pi_engine* e, unsigned char* table, const float* actions, int n_actions, const unsigned char* absorbing,
                 long long n_states, long long n_pad, long long s_begin) {
    // parameter block: natural alignment, arrays of PI_D entries — must match table_build_src.h
    std::vector<unsigned char> b;
    put(b, table, 8);
    put(b, actions, 8);
    put(b, absorbing, 8);
    for (int d = 0; d < e->D; ++d) { const float* ax = e->d_axes[d]; put(b, ax, 8); }
    put(b, n_states, 8);
    put(b, n_pad, 8);
    put(b, s_begin, 8);
    put(b, n_actions, 4);
    for (int d = 0; d < e->D; ++d) put(b, e->g.shape[d], 4);
    for (int d = 0; d < e->D; ++d) put(b, e->g.istride[d], 4);
    for (int d = 0; d < e->D; ++d) put(b, e->g.lo[d], 4);
    for (int d = 0; d < e->D; ++d) put(b, e->g.hi[d], 4);
    for (int d = 0; d < e->D; ++d) put(b, e->g.perm[d], 4);
    while (b.size() % 8) b.push_back(0);
    void* args[] = {b.data()};
    dim3 grid(nblocks(n_states), (unsigned)n_actions, 1);
    CU(cudaLaunchKernel((const void*)e->build_kernel, grid, dim3(pi::kBlock, 1, 1), args, 0, e->stream));
    e->launches++;
    return PI_OK;
}

// Layout probe: for every candidate fast dimension build the rows of the
// middle action for a few sample chunks and count the 128-byte lines a warp's V gather
// touches; keep the candidate with the fewest.  DPB200_FAST_DIM = ref | auto | <dim>.
int choose_layout(pi_engine* e) {
    const int D = e->D;
    const char* env = getenv("DPB200_FAST_DIM");
    if (env && (!strcmp(env, "ref") || !strcmp(env, "reference"))) { set_layout(e, D - 1); return PI_OK; }
    if (env && env[0] >= '0' && env[0] <= '9') {
        int f = atoi(env);
        if (f < 0 || f >= D) return fail(PI_ERR_INVALID, "DPB200_FAST_DIM=%d out of range", f);
        set_layout(e, f);
        return PI_OK;
    }
    const long long chunk = std::min<long long>(e->N, 16384);
    const int n_chunks = e->N > 8 * chunk ? 4 : 1;
    const long long n_pad = (chunk + 31) / 32 * 32;
    const size_t W = (size_t)D + 2;
    unsigned char* tmp = nullptr;
    unsigned long long* cnt = nullptr;
    CU(cudaMalloc(&tmp, W * 4 * (size_t)n_pad));
    CU(cudaMalloc(&cnt, 16));
    const int first_plane = W >= 4 ? 16 : (W >= 2 ? 8 : 4);
    int best = D - 1;
    double best_lines = 1e30;
    for (int f = D - 1; f >= 0; --f) {
        set_layout(e, f);
        CU(cudaMemsetAsync(cnt, 0, 16, e->stream));
        for (int c = 0; c < n_chunks; ++c) {
            const long long s0 = n_chunks == 1 ? 0 : (e->N / 8) * (2 * c + 1) / 32 * 32;
            const long long n = std::min(chunk, e->N - s0);
            int rc = launch_build(e, tmp, e->d_actions + e->A / 2, 1, nullptr, n, n_pad, s0);
            if (rc) return rc;
            pi::gather_lines_kernel<<<nblocks(n), pi::kBlock, 0, e->stream>>>(tmp, first_plane, n, cnt, cnt + 1);
            e->launches++;
        }
        unsigned long long h[2];
        CU(cudaMemcpyAsync(h, cnt, 16, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        const double lines = h[1] ? (double)h[0] / (double)h[1] : 32.0;
        e->probe_lines[f] = lines;
        if (lines < best_lines * 0.95) { best_lines = lines; best = f; }  // prefer the reference order on near-ties
    }
    cudaFree(tmp);
    cudaFree(cnt);
    set_layout(e, best);
    return PI_OK;
}

// -------------------------------------------------------------- dispatch ---
inline unsigned eval_blocks(const pi_engine* e) {
    return e->pair_kernel ? nblocks((e->n_local + 1) / 2) : nblocks(e->n_local);
}

template <int D, int F>
void launch_pair(pi_engine* e, const pi::EvalParams& p) {
    if constexpr (F < D && pi::corner_pos<D>(F) <= 2)
        pi::eval_sweep_pair_kernel<D, F><<<eval_blocks(e), pi::kBlock, 0, e->stream>>>(p);
}

template <int D>
bool pair_supported(int fast_dim) { return fast_dim >= 0 && fast_dim < D && pi::corner_pos<D>(fast_dim) <= 2; }

template <int D>
void launch_eval(pi_engine* e, int j, int check) {
    pi::EvalParams p{};
    p.partial = e->d_partial;
    p.j = j;
    p.check = check;
    p.lookahead = e->lookahead;
    p.rows = e->d_rows;
    p.V0 = e->d_V[0];
    p.V1 = e->d_V[1];
    p.ctl = e->d_ctl;
    p.n_local = e->n_local;
    p.n_pad = e->n_pad;
    p.s_begin = e->s_begin;
    p.gamma = e->cfg.gamma;
    for (int d = 0; d < D; ++d) p.stride[d] = e->g.istride[d];
    if (e->pair_kernel) {
        switch (e->fast_dim) {
            case 0: launch_pair<D, 0>(e, p); break;
            case 1: launch_pair<D, 1>(e, p); break;
            case 2: launch_pair<D, 2>(e, p); break;
            case 3: launch_pair<D, 3>(e, p); break;
            case 4: launch_pair<D, 4>(e, p); break;
            case 5: launch_pair<D, 5>(e, p); break;
            default: break;
        }
        return;
    }
    pi::eval_sweep_kernel<D><<<eval_blocks(e), pi::kBlock, 0, e->stream>>>(p);
}
template <int D>
void query_pair_supported(pi_engine* e, bool* out) { *out = pair_supported<D>(e->fast_dim); }
template <int D>
void query_eval_occupancy(pi_engine* e, int* blocks) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, pi::eval_sweep_kernel<D>, pi::kBlock, 0) != cudaSuccess)
        *blocks = 2;
}
template <int D>
void launch_improve(pi_engine* e) {
    pi::ImproveParams p{};
    p.table = e->d_table;
    p.rows = e->d_rows;
    p.V = e->d_V[e->cur];
    p.policy = e->d_policy;
    p.partial = reinterpret_cast<unsigned int*>(e->d_partial);
    p.n_local = e->n_local;
    p.n_pad = e->n_pad;
    p.n_actions = e->A;
    p.gamma = e->cfg.gamma;
    for (int d = 0; d < D; ++d) p.stride[d] = e->g.istride[d];
    pi::improve_kernel<D><<<nblocks(e->n_local), pi::kBlock, 0, e->stream>>>(p);
}
template <int D>
void launch_compact(pi_engine* e) {
    pi::compact_rows_kernel<D><<<nblocks(e->n_local), pi::kBlock, 0, e->stream>>>(e->d_table, e->d_rows, e->d_policy,
                                                                               e->n_local, e->n_pad);
}
template <int D>
void launch_expand(pi_engine* e, int action, long long s0, long long count, int* idx, float* w, float* rw,
                   unsigned char* tm) {
    const unsigned char* tab = e->d_table + (size_t)action * pi::row_table_bytes(D, e->n_pad);
    pi::expand_rows_kernel<D><<<nblocks(count), pi::kBlock, 0, e->stream>>>(tab, e->n_pad, s0, count, e->g, e->s_begin,
                                                                         e->n_local, idx, w, rw, tm);
}

#define DISPATCH_D(e, fn, ...)                                     \
    switch ((e)->D) {                                              \
        case 1: fn<1>(__VA_ARGS__); break;                         \
        case 2: fn<2>(__VA_ARGS__); break;                         \
        case 3: fn<3>(__VA_ARGS__); break;                         \
        case 4: fn<4>(__VA_ARGS__); break;                         \
        case 5: fn<5>(__VA_ARGS__); break;                         \
        case 6: fn<6>(__VA_ARGS__); break;                         \
        default: break;                                            \
    }

// ------------------------------------------------------- sharded exchange --
// After a sweep every rank holds fresh values only for its own slice of Vout.
// Each rank then receives, from every peer, the sub-range of that peer's slice
// that its transition rows actually reference (computed once after the table
// build; the degenerate case is an all-gather).  NCCL grouped send/recv over
// NVLink; captured into the sweep graphs.
int exchange_values(pi_engine* e, float* V) {
    if (e->world == 1) return PI_OK;
    NC(g_nccl.GroupStart());
    for (int r = 0; r < e->world; ++r) {
        if (r == e->rank) continue;
        if (e->give_hi[r] > e->give_lo[r])
            NC(g_nccl.Send(V + e->give_lo[r], (size_t)(e->give_hi[r] - e->give_lo[r]), ncclFloat32, r, e->comm, e->stream));
        if (e->need_hi[r] > e->need_lo[r])
            NC(g_nccl.Recv(V + e->need_lo[r], (size_t)(e->need_hi[r] - e->need_lo[r]), ncclFloat32, r, e->comm, e->stream));
    }
    NC(g_nccl.GroupEnd());
    return PI_OK;
}

// One batch = k sweeps (+ exchange each) + one bookkeeping kernel.
//   kBatchDecide : last sweep is a sync point -> reduce residual [+ all-reduce] + decide
//   kBatchPlain  : no residual at all, only the sweep counter advances
//   kBatchMeasure: residual of the last sweep is reduced but no convergence decision
enum { kBatchDecide = 0, kBatchPlain = 1, kBatchMeasure = 2 };

int enqueue_batch_raw(pi_engine* e, int k, int parity, int mode) {
    const int has_check = mode != kBatchPlain;
    for (int i = 0; i < k; ++i) {
        DISPATCH_D(e, launch_eval, e, i, (has_check && i == k - 1) ? 1 : 0);
        if (e->world > 1) {
            float* Vout = e->d_V[(parity + i + 1) & 1];
            int rc = exchange_values(e, Vout);
            if (rc) return rc;
        }
    }
    const int nb = (int)eval_blocks(e);
    const bool fused = (mode == kBatchDecide && e->world == 1);
    pi::eval_reduce_kernel<<<1, 1024, 0, e->stream>>>(e->d_ctl, e->d_partial, nb, k, has_check, fused ? 1 : 0,
                                                       e->cfg.theta);
    if (mode == kBatchDecide && e->world > 1) {
        NC(g_nccl.AllReduce(&e->d_ctl->last_delta, e->d_delta_g, 1, ncclFloat32, ncclMax, e->comm, e->stream));
        pi::eval_decide_kernel<<<1, 1, 0, e->stream>>>(e->d_ctl, e->d_delta_g, e->cfg.theta);
    }
    return PI_OK;
}

int build_graphs(pi_engine* e) {
    // [start parity][0: one sweep + decide, 1: sync sweeps + decide, 2: sync sweeps, counter only]
    for (int par = 0; par < 2; ++par) {
        for (int kind = 0; kind < 3; ++kind) {
            if (e->graphs[par][kind]) { cudaGraphExecDestroy(e->graphs[par][kind]); e->graphs[par][kind] = nullptr; }
            const int k = kind == 0 ? 1 : e->cfg.sync_interval;
            cudaGraph_t graph;
            CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_batch_raw(e, k, par, kind == 2 ? kBatchPlain : kBatchDecide);
            cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
            if (rc) return rc;
            if (ce != cudaSuccess) return fail(PI_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            CU(cudaGraphInstantiate(&e->graphs[par][kind], graph, 0));
            cudaGraphDestroy(graph);
        }
    }
    return PI_OK;
}

int enqueue_batch(pi_engine* e, int k, int parity, int mode) {
    const int sync = e->cfg.sync_interval;
    if (k == 1 && mode == kBatchDecide && e->graphs[parity][0]) {
        CU(cudaGraphLaunch(e->graphs[parity][0], e->stream));
    } else if (k == sync && mode == kBatchDecide && e->graphs[parity][1]) {
        CU(cudaGraphLaunch(e->graphs[parity][1], e->stream));
    } else if (k == sync && mode == kBatchPlain && e->graphs[parity][2]) {
        CU(cudaGraphLaunch(e->graphs[parity][2], e->stream));
    } else {
        int rc = enqueue_batch_raw(e, k, parity, mode);
        if (rc) return rc;
        CU(cudaGetLastError());
    }
    e->launches += k + 1;
    return PI_OK;
}

int compute_exchange_plan(pi_engine* e);

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* pi_last_error(void) { return g_err.c_str(); }
int pi_abi_version(void) { return PI_ABI_VERSION; }

int pi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pi_compile_check(const char* dynamics_src, int32_t n_dims, int64_t* cubin_bytes) {
    if (!dynamics_src) return fail(PI_ERR_INVALID, "null argument");
    if (n_dims < 1 || n_dims > PI_MAX_DIMS) return fail(PI_ERR_INVALID, "n_dims must be in [1,%d]", PI_MAX_DIMS);
    std::vector<char> cubin;
    std::string log;
    int rc = compile_to_cubin(n_dims, dynamics_src, cubin, log);
    if (rc) return rc;
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return PI_OK;
}

int pi_nccl_unique_id(uint8_t out[128]) {
    if (!out) return fail(PI_ERR_INVALID, "null argument");
    if (!g_nccl.load() || !g_nccl.GetUniqueId) return fail(PI_ERR_COMM, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, 128);
    return PI_OK;
}

int pi_create(const pi_grid* grid, const float* actions, int32_t n_actions, const pi_config* config,
              const char* dynamics_src, int32_t device, const pi_shard* shard, pi_engine** out) {
    if (!grid || !actions || !config || !dynamics_src || !out) return fail(PI_ERR_INVALID, "null argument");
    if (grid->n_dims < 1 || grid->n_dims > PI_MAX_DIMS) return fail(PI_ERR_INVALID, "n_dims must be in [1,%d]", PI_MAX_DIMS);
    if (n_actions < 1) return fail(PI_ERR_INVALID, "need at least one action");
    long long N = 1;
    for (int d = 0; d < grid->n_dims; ++d) {
        if (grid->shape[d] < 2) return fail(PI_ERR_INVALID, "every dimension needs >= 2 bins (dim %d has %d)", d, grid->shape[d]);
        if (!grid->axes[d]) return fail(PI_ERR_INVALID, "axes[%d] is null", d);
        N *= grid->shape[d];
        if (N > 0x7fffffffLL) return fail(PI_ERR_INVALID, "grid has more than 2^31-1 states (flat indices are int32, like the reference)");
    }
    int ndev = pi_device_count();
    if (ndev <= 0) return fail(PI_ERR_NO_DEVICE, "no CUDA device available: the B200 engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(PI_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);

    pi_engine* e = new pi_engine();
    e->D = grid->n_dims;
    e->A = n_actions;
    e->N = N;
    e->cfg = *config;
    if (e->cfg.sync_interval <= 0) e->cfg.sync_interval = 25;
    e->device = device;
    long long st = 1;
    e->g.n_dims = e->D;
    for (int d = e->D - 1; d >= 0; --d) {
        e->g.shape[d] = grid->shape[d];
        e->g.stride[d] = (int)st;
        e->g.lo[d] = grid->lo[d];
        e->g.hi[d] = grid->hi[d];
        st *= grid->shape[d];
    }
    set_layout(e, e->D - 1);
    if (shard && shard->world_size > 1) {
        e->rank = shard->rank;
        e->world = shard->world_size;
        if (e->rank < 0 || e->rank >= e->world) { delete e; return fail(PI_ERR_INVALID, "bad rank"); }
    }
    e->s_begin = (long long)e->rank * N / e->world;
    e->s_end = (long long)(e->rank + 1) * N / e->world;
    e->n_local = e->s_end - e->s_begin;
    e->n_pad = (e->n_local + 31) / 32 * 32;

#define CUX(call)                                                                                   \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            int _rc = fail(PI_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e));            \
            pi_destroy(e);                                                                          \
            return _rc;                                                                             \
        }                                                                                           \
    } while (0)

    CUX(cudaSetDevice(device));
    CUX(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) CUX(cudaEventCreateWithFlags(&e->ev[i], cudaEventDisableTiming));
    CUX(cudaEventCreate(&e->ev_t0));
    CUX(cudaEventCreate(&e->ev_t1));

    int rc = compile_builder(e, dynamics_src);
    if (rc) { pi_destroy(e); return rc; }

    const size_t W = (size_t)e->D + 2;
    // +32 B: the pair kernel reads aligned 64-bit windows that may end one element past N-1
    CUX(cudaMalloc(&e->d_V[0], (size_t)N * 4 + 32));
    CUX(cudaMalloc(&e->d_V[1], (size_t)N * 4 + 32));
    CUX(cudaMalloc(&e->d_policy, (size_t)e->n_pad * 4));
    CUX(cudaMalloc(&e->d_term, (size_t)e->n_pad));
    CUX(cudaMalloc(&e->d_table, (size_t)e->A * W * 4 * (size_t)e->n_pad));
    CUX(cudaMalloc(&e->d_rows, W * 4 * (size_t)e->n_pad));
    CUX(cudaMalloc(&e->d_actions, (size_t)e->A * 4));
    CUX(cudaMalloc(&e->d_ctl, sizeof(pi::Ctl)));
    CUX(cudaMalloc(&e->d_partial, (size_t)nblocks(e->n_local) * 4 + 4));  // >= any grid used below
    CUX(cudaMalloc(&e->d_delta_g, 4));
    CUX(cudaMalloc(&e->d_changed_g, 8));
    CUX(cudaMallocHost(&e->h_ctl, 4 * sizeof(pi::Ctl)));
    CUX(cudaMemsetAsync(e->d_V[0], 0, (size_t)N * 4 + 32, e->stream));
    CUX(cudaMemsetAsync(e->d_V[1], 0, (size_t)N * 4 + 32, e->stream));
    CUX(cudaMemsetAsync(e->d_policy, 0, (size_t)e->n_pad * 4, e->stream));
    CUX(cudaMemsetAsync(e->d_term, 0, (size_t)e->n_pad, e->stream));
    CUX(cudaMemsetAsync(e->d_ctl, 0, sizeof(pi::Ctl), e->stream));
    CUX(cudaMemcpyAsync(e->d_actions, actions, (size_t)e->A * 4, cudaMemcpyHostToDevice, e->stream));
    for (int d = 0; d < e->D; ++d) {
        CUX(cudaMalloc(&e->d_axes[d], (size_t)grid->shape[d] * 4));
        CUX(cudaMemcpyAsync(e->d_axes[d], grid->axes[d], (size_t)grid->shape[d] * 4, cudaMemcpyHostToDevice, e->stream));
    }
    CUX(cudaStreamSynchronize(e->stream));
    rc = choose_layout(e);
    if (rc) { pi_destroy(e); return rc; }
    {
        bool ok = false;
        DISPATCH_D(e, query_pair_supported, e, &ok);
        const char* kenv = getenv("DPB200_EVAL_KERNEL");  // scalar | pair (default: pair when supported)
        e->pair_kernel = ok && (kenv && !strcmp(kenv, "pair"));  // opt-in: the persistent scalar kernel is faster today
        cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
        DISPATCH_D(e, query_eval_occupancy, e, &e->blocks_per_sm);  // persistent grid = resident blocks
        e->lookahead = e->sm_count * e->blocks_per_sm;
        if (const char* b = getenv("DPB200_LOOKAHEAD")) e->lookahead = std::max(0, atoi(b));
    }

    if (e->world > 1) {
        if (!g_nccl.load()) { pi_destroy(e); return fail(PI_ERR_COMM, "libnccl.so.2 could not be loaded: %s", dlerror()); }
        ncclUniqueId id;
        memcpy(&id, shard->nccl_id, sizeof id);
        int r = g_nccl.CommInitRank(&e->comm, e->world, id, e->rank);
        if (r != 0) { pi_destroy(e); return fail(PI_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    }
#undef CUX
    *out = e;
    return PI_OK;
}

void pi_destroy(pi_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto& gp : e->graphs)
        for (auto& gx : gp)
            if (gx) cudaGraphExecDestroy(gx);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    cudaFree(e->d_V[0]); cudaFree(e->d_V[1]); cudaFree(e->d_policy); cudaFree(e->d_term);
    cudaFree(e->d_mask_full); cudaFree(e->d_scratch); cudaFree(e->d_table); cudaFree(e->d_rows); cudaFree(e->d_actions);
    cudaFree(e->d_ctl); cudaFree(e->d_partial); cudaFree(e->d_delta_g); cudaFree(e->d_changed_g);
    for (auto& a : e->d_axes) cudaFree(a);
    if (e->h_ctl) cudaFreeHost(e->h_ctl);
    for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
    if (e->ev_t0) cudaEventDestroy(e->ev_t0);
    if (e->ev_t1) cudaEventDestroy(e->ev_t1);
    if (e->lib) cudaLibraryUnload(e->lib);
    if (e->stream) cudaStreamDestroy(e->stream);
    cudaGetLastError();
    delete e;
}

int pi_set_log(pi_engine* e, pi_log_fn fn, void* user) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    e->log_fn = fn;
    e->log_user = user;
    return PI_OK;
}

static int upload_mask(pi_engine* e, const uint8_t* mask) {
    if (!e->d_mask_full) CU(cudaMalloc(&e->d_mask_full, (size_t)e->N));
    CU(cudaMemcpyAsync(e->d_mask_full, mask, (size_t)e->N, cudaMemcpyHostToDevice, e->stream));
    return PI_OK;
}

int pi_set_terminal(pi_engine* e, const uint8_t* mask, float value) {
    if (!e || !mask) return fail(PI_ERR_INVALID, "null argument");
    if (e->table_built) return fail(PI_ERR_INVALID, "pi_set_terminal must precede pi_build_table");
    CU(cudaSetDevice(e->device));
    int rc = upload_mask(e, mask);
    if (rc) return rc;
    pi::to_internal_kernel<unsigned char><<<nblocks(e->n_local), pi::kBlock, 0, e->stream>>>(e->g, e->d_mask_full, e->d_term,
                                                                                        e->s_begin, e->n_local);
    pi::fill_masked_kernel<<<nblocks(e->N), pi::kBlock, 0, e->stream>>>(e->g, e->d_V[0], e->d_V[1], e->d_mask_full, e->N, value);
    e->launches += 2;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return PI_OK;
}

int pi_set_values(pi_engine* e, const uint8_t* mask, float value) {
    if (!e || !mask) return fail(PI_ERR_INVALID, "null argument");
    CU(cudaSetDevice(e->device));
    int rc = upload_mask(e, mask);
    if (rc) return rc;
    pi::fill_masked_kernel<<<nblocks(e->N), pi::kBlock, 0, e->stream>>>(e->g, e->d_V[0], e->d_V[1], e->d_mask_full, e->N, value);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return PI_OK;
}

int pi_build_table(pi_engine* e) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaEventRecord(e->ev_t0, e->stream));
    {
        int rc = launch_build(e, e->d_table, e->d_actions, e->A, e->d_term, e->n_local, e->n_pad, e->s_begin);
        if (rc) return rc;
    }
    DISPATCH_D(e, launch_compact, e);
    e->launches++;
    CU(cudaEventRecord(e->ev_t1, e->stream));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
    e->stats.build_ms = ms;
    e->table_built = true;
    if (e->world > 1) {
        int rc = compute_exchange_plan(e);
        if (rc) return rc;
    }
    int rc = build_graphs(e);
    if (rc) return rc;
    logf(e, 1, "Transition table: %lld states x %d actions, %.1f MB, built in %.3f ms (storage: dim %d fastest)",
         e->n_local, e->A, (double)pi_table_bytes(e) / 1048576.0, ms, e->fast_dim);
    return PI_OK;
}

int pi_evaluate(pi_engine* e, float* delta_out, int32_t* sweeps_out) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (!e->table_built) return fail(PI_ERR_INVALID, "pi_build_table has not been called");
    CU(cudaSetDevice(e->device));
    const int max_eval = e->cfg.max_eval_iter;
    const int sync = e->cfg.sync_interval;
    const int cur0 = e->cur;
    pi::eval_begin_kernel<<<1, 1, 0, e->stream>>>(e->d_ctl, cur0);
    e->launches++;
    CU(cudaEventRecord(e->ev_t0, e->stream));

    int enq = 0;                 // sweeps enqueued so far
    int head = 0, tail = 0;      // ring of in-flight batches
    int check_idx[4] = {0, 0, 0, 0};
    pi::Ctl last{};
    last.check_delta = __builtin_inff();
    bool done = false;
    const int depth = 3;
    while (true) {
        while (!done && head - tail < depth && enq < max_eval) {
            int c = (enq % sync == 0) ? enq : (enq / sync + 1) * sync;  // next sync sweep index (:325)
            if (c > max_eval - 1) c = max_eval - 1;
            const int k = c - enq + 1;
            int rc = enqueue_batch(e, k, (cur0 + enq) & 1, kBatchDecide);
            if (rc) return rc;
            const int slot = head & 3;
            CU(cudaMemcpyAsync(&e->h_ctl[slot], e->d_ctl, sizeof(pi::Ctl), cudaMemcpyDeviceToHost, e->stream));
            CU(cudaEventRecord(e->ev[slot], e->stream));
            check_idx[slot] = c;
            enq += k;
            ++head;
        }
        if (head == tail) break;
        const int slot = tail & 3;
        CU(cudaEventSynchronize(e->ev[slot]));
        last = e->h_ctl[slot];
        ++tail;
        const int i = check_idx[slot];
        if (!done) {
            if (e->cfg.log_interval > 0 && i % e->cfg.log_interval == 0)
                logf(e, 0, "  Eval iter %5d | delta = %.4e", i, (double)last.check_delta);
            if (last.done) {
                done = true;
                logf(e, 2, "  Eval converged at iter %d | delta = %.2e", last.conv_sweep, (double)last.check_delta);
            }
        }
    }
    CU(cudaEventRecord(e->ev_t1, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
    e->stats.eval_ms += ms;
    e->cur = (cur0 + last.base) & 1;
    e->stats.eval_sweeps += last.base;
    e->stats.last_delta = last.check_delta;
    if (!done)
        logf(e, 3, "  Eval hit max_eval_iter=%d | delta = %.2e", max_eval, (double)last.check_delta);
    if (delta_out) *delta_out = last.check_delta;
    if (sweeps_out) *sweeps_out = last.base;
    e->results_gathered = false;
    return PI_OK;
}

int pi_improve(pi_engine* e, int32_t* stable, int64_t* n_changed) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (!e->table_built) return fail(PI_ERR_INVALID, "pi_build_table has not been called");
    CU(cudaSetDevice(e->device));
    CU(cudaEventRecord(e->ev_t0, e->stream));
    DISPATCH_D(e, launch_improve, e);
    pi::count_reduce_kernel<<<1, 1024, 0, e->stream>>>(e->d_ctl, reinterpret_cast<unsigned int*>(e->d_partial),
                                                        (int)nblocks(e->n_local));
    e->launches += 2;
    CU(cudaGetLastError());
    unsigned long long* src = &e->d_ctl->changed;
    if (e->world > 1) {
        NC(g_nccl.AllReduce(&e->d_ctl->changed, e->d_changed_g, 1, ncclUint64, ncclSum, e->comm, e->stream));
        src = e->d_changed_g;
    }
    CU(cudaEventRecord(e->ev_t1, e->stream));
    unsigned long long* h = reinterpret_cast<unsigned long long*>(&e->h_ctl[0]);
    CU(cudaMemcpyAsync(h, src, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
    e->stats.improve_ms += ms;
    e->stats.last_changed = (int64_t)*h;
    if (stable) *stable = (*h == 0) ? 1 : 0;
    if (n_changed) *n_changed = (int64_t)*h;
    e->results_gathered = false;
    return PI_OK;
}

int pi_run(pi_engine* e, pi_stats* stats) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (!e->table_built) {
        int rc = pi_build_table(e);
        if (rc) return rc;
    }
    e->stats.pi_iterations = 0;
    e->stats.converged = 0;
    for (int n = 0; n < e->cfg.max_pi_iter; ++n) {
        logf(e, 1, "-- PI Iteration %d/%d --", n + 1, e->cfg.max_pi_iter);
        float delta;
        int32_t sweeps;
        int rc = pi_evaluate(e, &delta, &sweeps);
        if (rc) return rc;
        int32_t stable;
        int64_t changed;
        rc = pi_improve(e, &stable, &changed);
        if (rc) return rc;
        e->stats.pi_iterations = n + 1;
        if (stable) {
            e->stats.converged = 1;
            logf(e, 2, "Policy Iteration converged at iteration %d.", n + 1);
            break;
        }
    }
    if (!e->stats.converged) logf(e, 3, "Policy Iteration hit max_pi_iter=%d.", e->cfg.max_pi_iter);
    if (stats) *stats = e->stats;
    return PI_OK;
}

// Make every rank's current V buffer complete (all ranks hold every slice).
static int gather_full_values(pi_engine* e) {
    if (e->world == 1 || e->results_gathered) return PI_OK;
    float* V = e->d_V[e->cur];
    NC(g_nccl.GroupStart());
    for (int r = 0; r < e->world; ++r) {
        if (r == e->rank) continue;
        const long long lo = (long long)r * e->N / e->world, hi = (long long)(r + 1) * e->N / e->world;
        NC(g_nccl.Send(V + e->s_begin, (size_t)e->n_local, ncclFloat32, r, e->comm, e->stream));
        NC(g_nccl.Recv(V + lo, (size_t)(hi - lo), ncclFloat32, r, e->comm, e->stream));
    }
    NC(g_nccl.GroupEnd());
    CU(cudaStreamSynchronize(e->stream));
    e->results_gathered = true;
    return PI_OK;
}

static int need_scratch(pi_engine* e) {
    if (!e->d_scratch) CU(cudaMalloc(&e->d_scratch, (size_t)e->N * 4));
    return PI_OK;
}

int pi_copy_results(pi_engine* e, float* value_function, int32_t* policy) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    CU(cudaSetDevice(e->device));
    if (!e->identity_layout) { int rc = need_scratch(e); if (rc) return rc; }
    if (value_function) {
        int rc = gather_full_values(e);
        if (rc) return rc;
        const float* src = e->d_V[e->cur];
        if (!e->identity_layout) {  // internal storage order -> reference order
            pi::to_reference_kernel<float><<<nblocks(e->N), pi::kBlock, 0, e->stream>>>(
                e->g, e->d_V[e->cur], static_cast<float*>(e->d_scratch), e->N);
            e->launches++;
            src = static_cast<float*>(e->d_scratch);
        }
        CU(cudaMemcpyAsync(value_function, src, (size_t)e->N * 4, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    if (policy) {
        const int* full = e->d_policy;  // single GPU: the local slice is everything
        if (e->world > 1) {
            // gather the policy slices through the spare V buffer (same element size)
            int* buf = reinterpret_cast<int*>(e->d_V[e->cur ^ 1]);
            CU(cudaMemcpyAsync(buf + e->s_begin, e->d_policy, (size_t)e->n_local * 4, cudaMemcpyDeviceToDevice, e->stream));
            NC(g_nccl.GroupStart());
            for (int r = 0; r < e->world; ++r) {
                if (r == e->rank) continue;
                const long long lo = (long long)r * e->N / e->world, hi = (long long)(r + 1) * e->N / e->world;
                NC(g_nccl.Send(buf + e->s_begin, (size_t)e->n_local, ncclInt32, r, e->comm, e->stream));
                NC(g_nccl.Recv(buf + lo, (size_t)(hi - lo), ncclInt32, r, e->comm, e->stream));
            }
            NC(g_nccl.GroupEnd());
            full = buf;
        }
        if (!e->identity_layout) {
            pi::to_reference_kernel<int><<<nblocks(e->N), pi::kBlock, 0, e->stream>>>(
                e->g, full, static_cast<int*>(e->d_scratch), e->N);
            e->launches++;
            full = static_cast<int*>(e->d_scratch);
        }
        CU(cudaMemcpyAsync(policy, full, (size_t)e->N * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    return PI_OK;
}

int pi_copy_local_results(pi_engine* e, float* v_local, int32_t* policy_local) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    CU(cudaSetDevice(e->device));
    if (v_local)
        CU(cudaMemcpyAsync(v_local, e->d_V[e->cur] + e->s_begin, (size_t)e->n_local * 4, cudaMemcpyDeviceToHost, e->stream));
    if (policy_local)
        CU(cudaMemcpyAsync(policy_local, e->d_policy, (size_t)e->n_local * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PI_OK;
}

int pi_upload_policy(pi_engine* e, const int32_t* policy) {
    if (!e || !policy) return fail(PI_ERR_INVALID, "null argument");
    if (!e->table_built) return fail(PI_ERR_INVALID, "pi_build_table has not been called");
    CU(cudaSetDevice(e->device));
    if (e->identity_layout) {
        CU(cudaMemcpyAsync(e->d_policy, policy + e->s_begin, (size_t)e->n_local * 4, cudaMemcpyHostToDevice, e->stream));
    } else {
        int rc = need_scratch(e);
        if (rc) return rc;
        CU(cudaMemcpyAsync(e->d_scratch, policy, (size_t)e->N * 4, cudaMemcpyHostToDevice, e->stream));
        pi::to_internal_kernel<int><<<nblocks(e->n_local), pi::kBlock, 0, e->stream>>>(
            e->g, static_cast<const int*>(e->d_scratch), e->d_policy, e->s_begin, e->n_local);
        e->launches++;
    }
    DISPATCH_D(e, launch_compact, e);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return PI_OK;
}

int pi_upload_values(pi_engine* e, const float* v) {
    if (!e || !v) return fail(PI_ERR_INVALID, "null argument");
    CU(cudaSetDevice(e->device));
    if (e->identity_layout) {
        CU(cudaMemcpyAsync(e->d_V[e->cur], v, (size_t)e->N * 4, cudaMemcpyHostToDevice, e->stream));
    } else {
        int rc = need_scratch(e);
        if (rc) return rc;
        CU(cudaMemcpyAsync(e->d_scratch, v, (size_t)e->N * 4, cudaMemcpyHostToDevice, e->stream));
        pi::to_internal_kernel<float><<<nblocks(e->N), pi::kBlock, 0, e->stream>>>(
            e->g, static_cast<const float*>(e->d_scratch), e->d_V[e->cur], 0, e->N);
        e->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(e->stream));
    e->results_gathered = true;
    return PI_OK;
}

int pi_sweeps(pi_engine* e, int32_t n_sweeps, float* delta, float* device_ms) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (!e->table_built) return fail(PI_ERR_INVALID, "pi_build_table has not been called");
    if (n_sweeps < 1) return fail(PI_ERR_INVALID, "n_sweeps must be >= 1");
    CU(cudaSetDevice(e->device));
    const int sync = e->cfg.sync_interval;
    const int cur0 = e->cur;
    pi::eval_begin_kernel<<<1, 1, 0, e->stream>>>(e->d_ctl, cur0);
    e->launches++;
    CU(cudaEventRecord(e->ev_t0, e->stream));
    int enq = 0;
    while (enq < n_sweeps) {
        const int k = std::min(sync, n_sweeps - enq);
        int rc = enqueue_batch(e, k, (cur0 + enq) & 1, enq + k == n_sweeps ? kBatchMeasure : kBatchPlain);
        if (rc) return rc;
        enq += k;
    }
    CU(cudaEventRecord(e->ev_t1, e->stream));
    CU(cudaMemcpyAsync(&e->h_ctl[0], e->d_ctl, sizeof(pi::Ctl), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
    e->cur = (cur0 + n_sweeps) & 1;
    e->stats.eval_sweeps += n_sweeps;
    e->stats.eval_ms += ms;
    if (delta) *delta = e->h_ctl[0].last_delta;
    if (device_ms) *device_ms = ms;
    e->results_gathered = false;
    return PI_OK;
}

int pi_expand_rows(pi_engine* e, int32_t action, int64_t s_begin, int64_t count, int32_t* idx, float* w,
                   float* reward, uint8_t* terminated) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (!e->table_built) return fail(PI_ERR_INVALID, "pi_build_table has not been called");
    if (action < 0 || action >= e->A) return fail(PI_ERR_INVALID, "action out of range");
    if (s_begin < 0 || s_begin + count > e->N || count < 0)
        return fail(PI_ERR_INVALID, "state range [%lld,%lld) outside the grid [0,%lld)", (long long)s_begin,
                    (long long)(s_begin + count), e->N);
    if (count == 0) return PI_OK;
    CU(cudaSetDevice(e->device));
    const int C = 1 << e->D;
    int* d_idx = nullptr; float* d_w = nullptr; float* d_r = nullptr; unsigned char* d_t = nullptr;
    if (idx) CU(cudaMalloc(&d_idx, (size_t)count * C * 4));
    if (w) CU(cudaMalloc(&d_w, (size_t)count * C * 4));
    if (reward) CU(cudaMalloc(&d_r, (size_t)count * 4));
    if (terminated) CU(cudaMalloc(&d_t, (size_t)count));
    DISPATCH_D(e, launch_expand, e, action, s_begin, count, d_idx, d_w, d_r, d_t);
    e->launches++;
    CU(cudaGetLastError());
    if (idx) CU(cudaMemcpyAsync(idx, d_idx, (size_t)count * C * 4, cudaMemcpyDeviceToHost, e->stream));
    if (w) CU(cudaMemcpyAsync(w, d_w, (size_t)count * C * 4, cudaMemcpyDeviceToHost, e->stream));
    if (reward) CU(cudaMemcpyAsync(reward, d_r, (size_t)count * 4, cudaMemcpyDeviceToHost, e->stream));
    if (terminated) CU(cudaMemcpyAsync(terminated, d_t, (size_t)count, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d_idx); cudaFree(d_w); cudaFree(d_r); cudaFree(d_t);
    return PI_OK;
}

int pi_device_ptrs(pi_engine* e, void** v, void** nv, void** policy, void** term) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (v) *v = e->d_V[e->cur];
    if (nv) *nv = e->d_V[e->cur ^ 1];
    if (policy) *policy = e->d_policy;
    if (term) *term = e->d_term;
    return PI_OK;
}

int pi_layout(const pi_engine* e, int32_t* fast_dim, int32_t perm[PI_MAX_DIMS], double probe_lines[PI_MAX_DIMS]) {
    if (!e) return fail(PI_ERR_INVALID, "null engine");
    if (fast_dim) *fast_dim = e->fast_dim;
    for (int d = 0; d < PI_MAX_DIMS; ++d) {
        if (perm) perm[d] = d < e->D ? e->g.perm[d] : -1;
        if (probe_lines) probe_lines[d] = d < e->D ? e->probe_lines[d] : 0.0;
    }
    return PI_OK;
}

int64_t pi_n_states(const pi_engine* e) { return e ? e->N : 0; }
int64_t pi_local_begin(const pi_engine* e) { return e ? e->s_begin : 0; }
int64_t pi_local_end(const pi_engine* e) { return e ? e->s_end : 0; }
int64_t pi_table_bytes(const pi_engine* e) {
    return e ? (int64_t)((size_t)(e->A + 1) * ((size_t)e->D + 2) * 4 * (size_t)e->n_pad) : 0;
}
int64_t pi_launch_count(const pi_engine* e) { return e ? e->launches : 0; }
int pi_get_stats(const pi_engine* e, pi_stats* stats) {
    if (!e || !stats) return fail(PI_ERR_INVALID, "null argument");
    *stats = e->stats;
    return PI_OK;
}

int pi_lookup_actions(pi_engine* e, const float* points, int64_t n_points, float* out) {
    (void)e; (void)points; (void)n_points; (void)out;
    return fail(PI_ERR_INVALID, "pi_lookup_actions: not implemented in this build");
}

}  // extern "C"

namespace {
// Needs-driven exchange plan: scan this rank's rows once (on the device) for the
// min / max flat index referenced inside every peer's slice.
__global__ void need_ranges_kernel(const unsigned char* table, long long n_pad, long long n_local, int n_actions,
                                   int row_words, int span, long long N, int world, long long* lo, long long* hi) {
    // base word is word 0 of plane 0: 16-byte planes when W>=4, else 8- or 4-byte plane.
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_local) return;
    const int first_plane_bytes = row_words >= 4 ? 16 : (row_words >= 2 ? 8 : 4);
    for (int a = 0; a < n_actions; ++a) {
        const unsigned char* tab = table + (size_t)a * ((size_t)row_words * 4u * (size_t)n_pad);
        const int base = *reinterpret_cast<const int*>(tab + (size_t)s * first_plane_bytes);
        if (base < 0) continue;
        const long long b0 = base, b1 = (long long)base + span;  // inclusive extent of the 2^D corners
        const int r0 = (int)(((b0 + 1) * world - 1) / N), r1 = (int)(((b1 + 1) * world - 1) / N);
        for (int r = r0; r <= r1 && r < world; ++r) {
            const long long rlo = (long long)r * N / world, rhi = (long long)(r + 1) * N / world;
            const long long a0 = b0 > rlo ? b0 : rlo, a1 = (b1 + 1) < rhi ? (b1 + 1) : rhi;
            if (a1 > a0) {
                atomicMin(reinterpret_cast<unsigned long long*>(&lo[r]), (unsigned long long)a0);
                atomicMax(reinterpret_cast<unsigned long long*>(&hi[r]), (unsigned long long)a1);
            }
        }
    }
}

int compute_exchange_plan(pi_engine* e) {
    const int Wd = e->world;
    std::vector<long long> lo(Wd, e->N), hi(Wd, 0);
    long long *d_lo, *d_hi;
    CU(cudaMalloc(&d_lo, Wd * 8));
    CU(cudaMalloc(&d_hi, Wd * 8));
    CU(cudaMemcpyAsync(d_lo, lo.data(), Wd * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_hi, hi.data(), Wd * 8, cudaMemcpyHostToDevice, e->stream));
    int span = 0;
    for (int d = 0; d < e->D; ++d) span += e->g.istride[d];
    need_ranges_kernel<<<nblocks(e->n_local), pi::kBlock, 0, e->stream>>>(e->d_table, e->n_pad, e->n_local, e->A,
                                                                        e->D + 2, span, e->N, Wd, d_lo, d_hi);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(lo.data(), d_lo, Wd * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(hi.data(), d_hi, Wd * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->need_lo.assign(Wd, 0);
    e->need_hi.assign(Wd, 0);
    for (int r = 0; r < Wd; ++r)
        if (r != e->rank && hi[r] > lo[r]) { e->need_lo[r] = lo[r]; e->need_hi[r] = hi[r]; }
    // tell every peer what I need from it: all-to-all of (lo,hi) pairs via NCCL send/recv
    std::vector<long long> mine(2 * Wd), theirs(2 * Wd, 0);
    for (int r = 0; r < Wd; ++r) { mine[2 * r] = e->need_lo[r]; mine[2 * r + 1] = e->need_hi[r]; }
    long long *d_m, *d_t;
    CU(cudaMalloc(&d_m, 2 * Wd * 8));
    CU(cudaMalloc(&d_t, 2 * Wd * 8));
    CU(cudaMemcpyAsync(d_m, mine.data(), 2 * Wd * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(d_t, 0, 2 * Wd * 8, e->stream));
    NC(g_nccl.GroupStart());
    for (int r = 0; r < Wd; ++r) {
        if (r == e->rank) continue;
        NC(g_nccl.Send(d_m + 2 * r, 2, ncclInt64, r, e->comm, e->stream));
        NC(g_nccl.Recv(d_t + 2 * r, 2, ncclInt64, r, e->comm, e->stream));
    }
    NC(g_nccl.GroupEnd());
    CU(cudaMemcpyAsync(theirs.data(), d_t, 2 * Wd * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->give_lo.assign(Wd, 0);
    e->give_hi.assign(Wd, 0);
    long long recv_total = 0;
    for (int r = 0; r < Wd; ++r) {
        e->give_lo[r] = theirs[2 * r];
        e->give_hi[r] = theirs[2 * r + 1];
        recv_total += e->need_hi[r] - e->need_lo[r];
    }
    cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_m); cudaFree(d_t);
    e->exchange_ready = true;
    logf(e, 1, "Shard %d/%d: states [%lld,%lld), receives %.1f MB of V per sweep (all-gather would be %.1f MB)",
         e->rank, Wd, e->s_begin, e->s_end, recv_total * 4.0 / 1048576.0, (e->N - e->n_local) * 4.0 / 1048576.0);
    return PI_OK;
}
}  // namespace
