// abi.cu — the extern "C" surface of librdb200.so (declared in include/rdb200.h): argument checking, model
// descriptors, host/device pointer classification, the pinned multi-stream host pipeline, and dispatch to the
// per-model kernel units (unit.cu) and the LieState kernels (lie.cu).  No torch, no Python, no CPU fallback:
// without a CUDA device every compute entry point fails with RDB_ERR_NO_DEVICE / a CUDA error code.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "rdb200.h"
#include "units.h"
#include "lie.h"
#include "custom.h"
#include "layout.h"
#include <map>
#include <vector>

using namespace rdb;

namespace {
constexpr int NSLOT = 3;                 // pipeline depth of the host-pointer path
constexpr long long HOST_CHUNK_DEFAULT = 1 << 16;  // knot points per pipelined chunk (RDB200_HOST_CHUNK overrides, for experiments)
enum { B_Z = 0, B_DT = 1, B_J = 2, B_OUT = 3, B_AUX = 4, B_T = 5, NBUF = 6 };

struct Slot {
    cudaStream_t st = nullptr;
    void* buf[NBUF] = {};
    size_t cap[NBUF] = {};
};
}  // namespace

// knot-major scratch for component-major callers, one set per stream that has used it
struct SoaScratch { void* buf[3] = {nullptr, nullptr, nullptr}; size_t cap[3] = {0, 0, 0}; };

struct rdb_context {
    int device = 0;
    int sm_count = 0;
    std::map<cudaStream_t, SoaScratch> soa;   // guarded by soa_mu
    std::mutex soa_mu;
    int pdl = 1;     // programmatic dependent launch of the knot kernels (RDB200_PDL=0 disables)
    long long host_chunk = HOST_CHUNK_DEFAULT;
    Slot slot[NSLOT];
    cudaEvent_t ev_in = nullptr;   // device-resident inputs + host outputs: the slot streams wait for the caller's stream here
    std::mutex mu;   // the staging slots are shared by all host-pointer calls on this context
};

struct rdb_model {
    rdb_context* ctx;
    int kind, rot, frame, D;
    int n, m, nerr;
    ModelParams<double> p;
    CustomModel* custom;   // kind == RDB_CUSTOM: NVRTC-compiled user model (custom.cu)
    LieParts lie;          // LieState{R,P} partition of the state (Euclidean: one vector block; RigidBody{R}: (3, 6))
    int general_lie;       // user model with an arbitrary LieState{R,P} (rdb_model_create_custom_lie): kernels see a Euclidean model,
                           // the LieState maps and the error-state projection use `lie`
};

// Persistent device mirror of a batch of SampledTrajectories (reference: src/trajectories.jl:40-50), knot-major across the batch:
// row k * ntraj + j of Z holds z = [x;u] of knot k of trajectory j; t and dt are stored the same way.  With this order the knots of a
// time chunk [kb, ke] of ALL trajectories are one contiguous range (what the Jacobian kernel streams), a warp of the rollout kernel
// (adjacent trajectories) reads and writes one contiguous range per step, and for ntraj == 1 the image is exactly the gathered
// Matrix(n+m, K) of a Vector{KnotPoint}.
struct rdb_trajectory {
    const rdb_model* M = nullptr;
    int dtype = RDB_F64;
    long long ntraj = 0;
    int K = 0;
    void* Z = nullptr;
    double* t = nullptr;
    double* dt = nullptr;
    void* stage = nullptr; size_t stage_cap = 0;        // dense device staging of the host-pointer setters / getters
    cudaStream_t aux[2] = {nullptr, nullptr};           // rollout / linearisation pipeline of rdb_trajectory_rollout_linearize
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_chunk;
    std::mutex mu;
};

// A validated, pre-dispatched knot operation on device pointers: launching it is one kernel launch and nothing else.
struct rdb_plan {
    const rdb_model* M = nullptr;
    int dtype = RDB_F64, layout = RDB_AOS;
    KnotRequest r;
};

namespace {

int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : int(e); }

// Every entry point runs on its context's device and puts the caller's current device back before returning: in a multi-GPU
// process (one torch process driving several devices) the library must not change where the caller's next allocation lands.
struct DeviceGuard {
    int prev = -1; bool changed = false; cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != d) { err = cudaSetDevice(d); changed = (err == cudaSuccess) && prev >= 0; }
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define RDB_ON_DEVICE(ctx) DeviceGuard guard__((ctx)->device); if (guard__.err != cudaSuccess) return int(guard__.err)
#define RDB_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return int(e__); } while (0)

void inv3(const double* A, double* Ai) {
    const double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], h = A[7], i = A[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    Ai[0] = (e * i - f * h) * id; Ai[1] = (c * h - b * i) * id; Ai[2] = (b * f - c * e) * id;
    Ai[3] = (f * g - d * i) * id; Ai[4] = (a * i - c * g) * id; Ai[5] = (c * d - a * f) * id;
    Ai[6] = (d * h - e * g) * id; Ai[7] = (b * g - a * h) * id; Ai[8] = (a * e - b * d) * id;
}

// 0 = null, 1 = host (pinned, registered or pageable), 2 = device / managed, RDB_ERR_POINTER_MIX = memory of ANOTHER device
int ptr_kind(const void* p, int device = -1) {
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 1; }
    if (at.type == cudaMemoryTypeDevice && device >= 0 && at.device != device) return RDB_ERR_POINTER_MIX;
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? 2 : 1;
}
// classify a set of data pointers: returns 1 host, 2 device, 0 all-null, RDB_ERR_POINTER_MIX on a mix (or on device memory that
// does not belong to `device`, the context's GPU)
int classify(std::initializer_list<const void*> ps, int device = -1) {
    int kind = 0;
    for (const void* p : ps) {
        const int k = ptr_kind(p, device);
        if (k == 0) continue;
        if (k < 0) return k;
        if (kind == 0) kind = k;
        else if (kind != k) return RDB_ERR_POINTER_MIX;
    }
    return kind;
}

int ensure(Slot& s, int i, size_t bytes) {
    if (bytes <= s.cap[i]) return 0;
    if (s.buf[i]) { RDB_CUDA(cudaFree(s.buf[i])); s.buf[i] = nullptr; s.cap[i] = 0; }
    RDB_CUDA(cudaMalloc(&s.buf[i], bytes));
    s.cap[i] = bytes;
    return 0;
}

size_t esize(int dtype) { return dtype == RDB_F32 ? 4 : 8; }

// copy `rows` component streams of `cnt` knots between a host array (leading dimension N) and a packed device chunk
int copy_chunk(void* dst, const void* src, int layout, size_t es, int width, long long N, long long k0, long long cnt,
               cudaMemcpyKind dir, cudaStream_t st) {
    if (layout == RDB_AOS) {
        const size_t off = size_t(k0) * width * es, bytes = size_t(cnt) * width * es;
        if (dir == cudaMemcpyHostToDevice) return cuda_rc(cudaMemcpyAsync(dst, (const char*)src + off, bytes, dir, st));
        return cuda_rc(cudaMemcpyAsync((char*)dst + off, src, bytes, dir, st));
    }
    if (dir == cudaMemcpyHostToDevice)
        return cuda_rc(cudaMemcpy2DAsync(dst, size_t(cnt) * es, (const char*)src + size_t(k0) * es, size_t(N) * es, size_t(cnt) * es, width, dir, st));
    return cuda_rc(cudaMemcpy2DAsync((char*)dst + size_t(k0) * es, size_t(N) * es, src, size_t(cnt) * es, size_t(cnt) * es, width, dir, st));
}

int sync_slots(rdb_context* c) {
    int rc = 0;
    for (auto& s : c->slot) { const int r = cuda_rc(cudaStreamSynchronize(s.st)); if (r && !rc) rc = r; }
    return rc;
}

int map_q(int integrator) {
    switch (integrator) {
        case RDB_EULER: return Q_EULER; case RDB_RK2: return Q_RK2; case RDB_RK3: return Q_RK3; case RDB_RK4: return Q_RK4;
        case RDB_IMPLICIT_MIDPOINT: return Q_IMPLICIT_MIDPOINT;
    }
    return -1;
}

// built-in models go to their compilation unit, user models to the NVRTC path
int dispatch(const rdb_model* M, int dtype, KnotRequest* r) {
    if (M->kind == RDB_CUSTOM) return custom_run(M->custom, *r);
    UnitFn fn = find_unit(M->kind, M->rot, M->frame, M->D, dtype);
    return fn ? fn(r) : RDB_ERR_NOT_IMPLEMENTED;
}

// Component-major ("SoA") callers: transpose a chunk of knots to the knot-major layout the kernels stream, evaluate, transpose the
// results back.  All on the caller's stream; scratch is per (context, stream), so calls on one stream simply queue up.
constexpr size_t SOA_SCRATCH_BYTES = size_t(256) << 20;     // knot-major Jacobian scratch per stream: bounds the chunk length
int dispatch_soa(const rdb_model* M, int dtype, const KnotRequest& r, long long ld) {
    rdb_context* c = M->ctx;
    const size_t es = esize(dtype);
    const int n = M->n, NZ = M->n + M->m, E = r.err ? M->nerr * (M->nerr + M->m) : n * NZ;
    // Built-in models: the kernel itself reads and writes component-major arrays through 2-D tensor maps (kernels.cuh, SOA = true)
    // when the component rows are whole 16-byte units; everything else (odd row lengths, unaligned pointers, user models,
    // ImplicitMidpoint) is transposed around the knot-major kernel.
    if (M->kind != RDB_CUSTOM && r.Q != Q_IMPLICIT_MIDPOINT && soa_tma_ok(r.Z, r.J, r.out, ld, int(es))) {
        KnotRequest q = r;
        q.soa = 1; q.ld = ld;
        const int rc = dispatch(M, dtype, &q);
        if (rc != RDB_ERR_NOT_IMPLEMENTED) return rc;
    }
    // the scratch of a stream is shared by every host thread that submits on it (notably the NULL stream): the lock is held for the
    // whole enqueue sequence, so two threads cannot interleave transpose / kernel / transpose on the same buffers
    std::lock_guard<std::mutex> lock(c->soa_mu);
    SoaScratch* sc = &c->soa[r.stream];
    long long SOA_CHUNK = (long long)(SOA_SCRATCH_BYTES / (size_t(E) * es));
    SOA_CHUNK = SOA_CHUNK < 1024 ? 1024 : (SOA_CHUNK / 1024) * 1024;
    const long long cap = r.N < SOA_CHUNK ? r.N : SOA_CHUNK;
    const size_t need[3] = {size_t(cap) * NZ * es, r.J ? size_t(cap) * E * es : 0, r.out ? size_t(cap) * n * es : 0};
    for (int i = 0; i < 3; ++i)
        if (need[i] > sc->cap[i]) {
            if (sc->buf[i]) RDB_CUDA(cudaFree(sc->buf[i]));
            sc->buf[i] = nullptr; sc->cap[i] = 0;
            RDB_CUDA(cudaMalloc(&sc->buf[i], need[i]));
            sc->cap[i] = need[i];
        }
    for (long long k0 = 0; k0 < r.N; k0 += SOA_CHUNK) {
        const long long cnt = (r.N - k0 < SOA_CHUNK) ? (r.N - k0) : SOA_CHUNK;
        int rc = soa_to_aos(dtype, (const char*)r.Z + size_t(k0) * es, ld, sc->buf[0], NZ, cnt, r.stream);
        if (rc) return rc;
        KnotRequest q = r;
        q.Z = sc->buf[0]; q.J = r.J ? sc->buf[1] : nullptr; q.out = r.out ? sc->buf[2] : nullptr; q.N = cnt;
        q.dt = r.dt ? r.dt + k0 : nullptr;
        q.t = r.t ? r.t + k0 : nullptr;
        if ((rc = dispatch(M, dtype, &q))) return rc;
        if (r.J && (rc = aos_to_soa(dtype, sc->buf[1], (char*)r.J + size_t(k0) * es, ld, E, cnt, r.stream))) return rc;
        if (r.out && (rc = aos_to_soa(dtype, sc->buf[2], (char*)r.out + size_t(k0) * es, ld, n, cnt, r.stream))) return rc;
    }
    return 0;
}

// KnotPoint.t reaches only models whose dynamics can depend on it: user models (dynamics(model, x, u, t), src/dynamics.jl:81-83).
// The shipped families are time-invariant, so their kernels never load it and the host path never copies it.
bool model_uses_time(const rdb_model* M) { return M->kind == RDB_CUSTOM; }

int knot_op(const rdb_model* M, int Q, int dtype, int layout, int with_j, long long N, const void* Z, const double* t, const double* dt,
            double dt0, void* J, void* out, void* stream, int err = 0);

// Error-state Jacobian of a user model with an arbitrary LieState{R,P}: the full-state Jacobian and x+ go to stream-ordered scratch,
// then one projection kernel forms G(x+)' [A B] blkdiag(G(x), I) (lie.cu).  Two passes: the G-seeded one-pass form of the rigid bodies is
// specialised to LieState(R, (3, 6)).
int general_error_jacobian(const rdb_model* M, int Q, int dtype, int layout, long long N, const void* Z, const double* t, const double* dt,
                           double dt0, void* Jbar, void* xn, void* stream) {
    if (layout != RDB_AOS) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    const int kind = classify({Z, t, dt, Jbar, xn}, c->device);
    if (kind < 0) return kind;
    const size_t es = esize(dtype);
    const int n = M->n, NZ = M->n + M->m, ne = M->nerr;
    cudaStream_t st = kind == 2 ? (cudaStream_t)stream : c->slot[0].st;
    void *dJ = nullptr, *dX = nullptr, *dZ = nullptr, *dB = nullptr;
    double *dT = nullptr, *dDt = nullptr;
    int rc = 0;
    auto alloc = [&](void** p, size_t bytes) { if (!rc) rc = cuda_rc(cudaMallocAsync(p, bytes, st)); };
    alloc(&dJ, size_t(N) * n * NZ * es);
    alloc(&dX, size_t(N) * n * es);
    const void* Zd = Z; const double* td = t; const double* dtd = dt; void* Bd = Jbar;
    if (kind == 1) {
        alloc(&dZ, size_t(N) * NZ * es); alloc(&dB, size_t(N) * ne * (ne + M->m) * es);
        if (!rc) rc = cuda_rc(cudaMemcpyAsync(dZ, Z, size_t(N) * NZ * es, cudaMemcpyHostToDevice, st));
        if (t && model_uses_time(M)) { alloc((void**)&dT, size_t(N) * 8); if (!rc) rc = cuda_rc(cudaMemcpyAsync(dT, t, size_t(N) * 8, cudaMemcpyHostToDevice, st)); }
        if (dt) { alloc((void**)&dDt, size_t(N) * 8); if (!rc) rc = cuda_rc(cudaMemcpyAsync(dDt, dt, size_t(N) * 8, cudaMemcpyHostToDevice, st)); }
        Zd = dZ; td = dT; dtd = dDt; Bd = dB;
    }
    if (!rc) rc = knot_op(M, Q, dtype, RDB_AOS, 1, N, Zd, td, dtd, dt0, dJ, dX, st, 0);
    if (!rc) rc = lie_project_error_jacobian(dtype, M->rot, M->lie, n, M->m, ne, N, Zd, NZ, dX, dJ, Bd, c->sm_count, st);
    if (!rc && kind == 1) rc = cuda_rc(cudaMemcpyAsync(Jbar, dB, size_t(N) * ne * (ne + M->m) * es, cudaMemcpyDeviceToHost, st));
    if (!rc && xn) rc = cuda_rc(cudaMemcpyAsync(xn, dX, size_t(N) * n * es, kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    for (void* p : {dJ, dX, dZ, dB, (void*)dT, (void*)dDt}) if (p) cudaFreeAsync(p, st);
    if (kind == 1) { const int r2 = cuda_rc(cudaStreamSynchronize(st)); if (!rc) rc = r2; }
    return rc;
}

// The one knot-point operation behind rdb_dynamics / rdb_discrete_dynamics / rdb_jacobian / rdb_discrete_jacobian.
int knot_op(const rdb_model* M, int Q, int dtype, int layout, int with_j, long long N, const void* Z, const double* t, const double* dt,
            double dt0, void* J, void* out, void* stream, int err) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || (layout != RDB_AOS && layout != RDB_SOA)) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!Z || (with_j && !J) || (!with_j && !out)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    if (!model_uses_time(M)) t = nullptr;
    // inputs and outputs are classified separately: device-resident inputs (a persistent device trajectory) may be combined with
    // HOST outputs (the solver keeps its Jacobians on the host) — the kernel then reads Z in place and only J / x+ cross PCIe.
    const int kind_in = classify({Z, t, dt}, c->device), kind_out = classify({J, out}, c->device);
    if (kind_in < 0) return kind_in;
    if (kind_out < 0) return kind_out;
    const bool dev_in_host_out = (kind_in == 2 && kind_out == 1);
    if (kind_in != kind_out && !dev_in_host_out) return RDB_ERR_POINTER_MIX;
    if (dev_in_host_out && layout != RDB_AOS) return RDB_ERR_POINTER_MIX;
    const int kind = kind_out;

    KnotRequest r;
    std::memset(&r, 0, sizeof(r));
    if (M->rot == RDB_ROT_NONE) err = 0;     // EuclideanState: G = I (src/statevectortype.jl:149-155)
    if (err && M->general_lie) return general_error_jacobian(M, Q, dtype, layout, N, Z, t, dt, dt0, J, out, stream);
    r.op = OP_KNOT; r.Q = Q; r.dtype = dtype; r.with_j = with_j; r.err = err; r.params = M->p;
    r.dt0 = dt0; r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
    if (kind == 2) {
        r.Z = Z; r.dt = dt; r.t = t; r.J = J; r.out = out; r.N = N; r.stream = (cudaStream_t)stream;
        return layout == RDB_SOA ? dispatch_soa(M, dtype, r, N) : dispatch(M, dtype, &r);
    }
    // host pointers: H2D -> kernel -> D2H per chunk, chunks round-robin over NSLOT streams so the three overlap
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    const int n = M->n, NZ = M->n + M->m, E = err ? M->nerr * (M->nerr + M->m) : n * NZ;
    int rc = 0;
    long long ci = 0;
    const long long HOST_CHUNK = c->host_chunk;
    const long long cap_knots = N < HOST_CHUNK ? N : HOST_CHUNK;
    if (dev_in_host_out) {      // the inputs are produced on the caller's stream: the slot streams start after it
        if (!c->ev_in) RDB_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
        RDB_CUDA(cudaEventRecord(c->ev_in, (cudaStream_t)stream));
        for (auto& s : c->slot) RDB_CUDA(cudaStreamWaitEvent(s.st, c->ev_in, 0));
    }
    for (long long k0 = 0; k0 < N && !rc; k0 += HOST_CHUNK, ++ci) {
        const long long cnt = (N - k0 < HOST_CHUNK) ? (N - k0) : HOST_CHUNK;
        Slot& s = c->slot[ci % NSLOT];
        if (J && (rc = ensure(s, B_J, size_t(cap_knots) * E * es))) break;
        if (out && (rc = ensure(s, B_OUT, size_t(cap_knots) * n * es))) break;
        if (dev_in_host_out) {
            r.Z = (const char*)Z + size_t(k0) * NZ * es; r.dt = dt ? dt + k0 : nullptr; r.t = t ? t + k0 : nullptr;
        } else {
            if ((rc = ensure(s, B_Z, size_t(cap_knots) * NZ * es))) break;
            if (dt && (rc = ensure(s, B_DT, size_t(cap_knots) * 8))) break;
            if (t && (rc = ensure(s, B_T, size_t(cap_knots) * 8))) break;
            if ((rc = copy_chunk(s.buf[B_Z], Z, layout, es, NZ, N, k0, cnt, cudaMemcpyHostToDevice, s.st))) break;
            if (dt && (rc = cuda_rc(cudaMemcpyAsync(s.buf[B_DT], dt + k0, size_t(cnt) * 8, cudaMemcpyHostToDevice, s.st)))) break;
            if (t && (rc = cuda_rc(cudaMemcpyAsync(s.buf[B_T], t + k0, size_t(cnt) * 8, cudaMemcpyHostToDevice, s.st)))) break;
            r.Z = s.buf[B_Z]; r.dt = dt ? (const double*)s.buf[B_DT] : nullptr; r.t = t ? (const double*)s.buf[B_T] : nullptr;
        }
        r.J = J ? s.buf[B_J] : nullptr; r.out = out ? s.buf[B_OUT] : nullptr; r.N = cnt; r.stream = s.st;
        if ((rc = (layout == RDB_SOA ? dispatch_soa(M, dtype, r, cnt) : dispatch(M, dtype, &r)))) break;
        if (J && (rc = copy_chunk(J, s.buf[B_J], layout, es, E, N, k0, cnt, cudaMemcpyDeviceToHost, s.st))) break;
        if (out && (rc = copy_chunk(out, s.buf[B_OUT], layout, es, n, N, k0, cnt, cudaMemcpyDeviceToHost, s.st))) break;
    }
    const int rs = sync_slots(c);
    return rc ? rc : rs;
}

// host staging for the small LieState / rollout operations: whole arrays through slot 0
struct Staged {
    Slot& s; int rc = 0;
    explicit Staged(Slot& slot) : s(slot) {}
    void* in(int b, const void* host, size_t bytes) {
        if (rc || !host) return nullptr;
        if ((rc = ensure(s, b, bytes))) return nullptr;
        rc = cuda_rc(cudaMemcpyAsync(s.buf[b], host, bytes, cudaMemcpyHostToDevice, s.st));
        return s.buf[b];
    }
    void* outbuf(int b, size_t bytes) { if (rc) return nullptr; rc = ensure(s, b, bytes); return s.buf[b]; }
    void back(void* host, int b, size_t bytes) { if (!rc) rc = cuda_rc(cudaMemcpyAsync(host, s.buf[b], bytes, cudaMemcpyDeviceToHost, s.st)); }
    int finish() { const int r = cuda_rc(cudaStreamSynchronize(s.st)); return rc ? rc : r; }
};

}  // namespace

extern "C" {

int rdb_version(void) { return 100; }

const char* rdb_strerror(int code) {
    switch (code) {
        case RDB_OK: return "ok";
        case RDB_ERR_ARG: return "invalid argument (bad enum, NULL pointer or negative size)";
        case RDB_ERR_NOT_IMPLEMENTED: return "not implemented for this model / integrator / dtype";
        case RDB_ERR_POINTER_MIX: return "host and device data pointers mixed in one call";
        case RDB_ERR_NO_DEVICE: return "no CUDA device available";
        case RDB_ERR_COMPILE: return "user model failed to compile (see rdb_last_log())";
    }
    if (code > 0) return cudaGetErrorString(cudaError_t(code));
    return "unknown rdb200 status";
}

int rdb_create(int device, rdb_context** ctx) {
    if (!ctx) return RDB_ERR_ARG;
    *ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return RDB_ERR_NO_DEVICE; }
    if (device < 0 || device >= count) return RDB_ERR_ARG;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return int(guard.err);
    rdb_context* c = new (std::nothrow) rdb_context();
    if (!c) return RDB_ERR_ARG;
    c->device = device;
    cudaError_t e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char* v = std::getenv("RDB200_PDL")) c->pdl = (v[0] != '0');
    if (const char* v = std::getenv("RDB200_HOST_CHUNK")) { const long long n = std::atoll(v); if (n >= 1024) c->host_chunk = n; }
    for (auto& s : c->slot) if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { rdb_destroy(c); return int(e); }       // frees the streams created so far and the context
    *ctx = c;
    return 0;
}

int rdb_destroy(rdb_context* c) {
    if (!c) return 0;
    DeviceGuard guard(c->device);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    for (auto& s : c->slot) {
        if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
        for (auto& b : s.buf) if (b) cudaFree(b);
    }
    cudaDeviceSynchronize();
    for (auto& kv : c->soa) for (auto& b : kv.second.buf) if (b) cudaFree(b);
    delete c;
    return 0;
}

void* rdb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void rdb_host_free(void* p) { if (p) cudaFreeHost(p); }
int rdb_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return RDB_ERR_ARG;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return int(e); }
    return 0;
}
int rdb_host_unregister(void* p) {
    if (!p) return RDB_ERR_ARG;
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return int(e); }
    return 0;
}

int rdb_model_create(rdb_context* ctx, int kind, int rot, int frame, const double* params, int np, rdb_model** model) {
    if (!ctx || !model || !params || np < 0) return RDB_ERR_ARG;
    *model = nullptr;
    rdb_model M;
    std::memset(&M, 0, sizeof(M));
    M.ctx = ctx; M.kind = kind; M.rot = rot; M.frame = frame; M.D = 0;
    ModelParams<double>& p = M.p;
    const int nrot = (rot == RDB_ROT_QUAT) ? 4 : 3;
    switch (kind) {
        case RDB_CARTPOLE:
            if (np < 4) return RDB_ERR_ARG;
            p.mc = params[0]; p.mp = params[1]; p.l = params[2]; p.g = params[3];
            p.cp_ia = 1.0 / (p.mp * p.l); p.cp_H00 = (p.mc + p.mp) * p.cp_ia; p.cp_nH00i = -1.0 / p.cp_H00;
            M.rot = RDB_ROT_NONE; M.frame = 0; M.n = 4; M.m = 1; M.nerr = 4;
            break;
        case RDB_QUADROTOR:
        case RDB_BODY: {
            const int need = (kind == RDB_QUADROTOR) ? 16 : 10;
            if (np < need || rot < RDB_ROT_QUAT || rot > RDB_ROT_RP || (frame != 0 && frame != 1)) return RDB_ERR_ARG;
            p.mass = params[0]; p.inv_mass = 1.0 / params[0];
            for (int i = 0; i < 9; ++i) p.J[i] = params[1 + i];
            inv3(p.J, p.Jinv);
            if (kind == RDB_QUADROTOR) {
                for (int i = 0; i < 3; ++i) p.mg[i] = p.mass * params[10 + i];
                p.motor_dist = params[13]; p.kf = params[14]; p.km = params[15];
            }
            M.n = 9 + nrot; M.m = (kind == RDB_QUADROTOR) ? 4 : 6; M.nerr = 12;
            break;
        }
        case RDB_DOUBLE_INTEGRATOR:
            if (np < 1) return RDB_ERR_ARG;
            M.D = int(params[0]);
            if (M.D < 1 || M.D > 3) return RDB_ERR_NOT_IMPLEMENTED;
            M.rot = RDB_ROT_NONE; M.frame = 0; M.n = 2 * M.D; M.m = M.D; M.nerr = M.n;
            break;
        default: return RDB_ERR_ARG;
    }
    M.lie = (M.rot == RDB_ROT_NONE) ? lie_parts_euclidean(M.n) : lie_parts_rigid();
    rdb_model* out = new (std::nothrow) rdb_model(M);
    if (!out) return RDB_ERR_ARG;
    *model = out;
    return 0;
}

int rdb_model_create_custom(rdb_context* ctx, int n, int m, const char* f_body, const double* params, int np, rdb_model** model) {
    if (!ctx || !model || !f_body || n < 1 || m < 1 || n + m > 32 || np < 0 || (np > 0 && !params)) return RDB_ERR_ARG;
    *model = nullptr;
    // fail at creation, with a readable log, rather than at the first evaluation
    if (custom_check(n, m, f_body, np, RDB_F64) != 0) return RDB_ERR_COMPILE;
    rdb_model M;
    std::memset(&M, 0, sizeof(M));
    M.ctx = ctx; M.kind = RDB_CUSTOM; M.rot = RDB_ROT_NONE; M.n = n; M.m = m; M.nerr = n;
    M.lie = lie_parts_euclidean(n);
    M.custom = custom_create(n, m, f_body, params, np);
    if (!M.custom) return RDB_ERR_ARG;
    *model = new (std::nothrow) rdb_model(M);
    return *model ? 0 : RDB_ERR_ARG;
}

int rdb_model_create_custom_lie(rdb_context* ctx, int rot, int nparts, const int* parts, int m, const char* f_body, const double* params, int np,
                                rdb_model** model) {
    if (!ctx || !model || !f_body || !parts || nparts < 1 || nparts > 8 || m < 1 || np < 0 || (np > 0 && !params) || rot < RDB_ROT_QUAT || rot > RDB_ROT_RP)
        return RDB_ERR_ARG;
    *model = nullptr;
    LieParts lp{};
    lp.nv = nparts;
    int nvec = 0;
    for (int i = 0; i < nparts; ++i) { if (parts[i] < 0) return RDB_ERR_ARG; lp.P[i] = parts[i]; nvec += parts[i]; }
    const int nrot = nparts - 1, w = (rot == RDB_ROT_QUAT) ? 4 : 3;
    const int n = nvec + nrot * w, ne = nvec + 3 * nrot;                 // length(LieState), errstate_dim (src/liestate.jl:118-124)
    if (n < 1 || n + m > 32) return RDB_ERR_ARG;
    if (custom_check(n, m, f_body, np, RDB_F64) != 0) return RDB_ERR_COMPILE;
    rdb_model M;
    std::memset(&M, 0, sizeof(M));
    M.ctx = ctx; M.kind = RDB_CUSTOM; M.rot = rot; M.n = n; M.m = m; M.nerr = ne; M.lie = lp; M.general_lie = 1;
    M.custom = custom_create(n, m, f_body, params, np);                  // the kernels see a plain (Euclidean) user model
    if (!M.custom) return RDB_ERR_ARG;
    *model = new (std::nothrow) rdb_model(M);
    return *model ? 0 : RDB_ERR_ARG;
}

int rdb_model_create_custom_rigid(rdb_context* ctx, int rot, int frame, int m, const char* wrench_body, double mass, const double* J,
                                  const double* params, int np, rdb_model** model) {
    if (!ctx || !model || !wrench_body || !J || m < 1 || m > 12 || np < 0 || (np > 0 && !params) || rot < RDB_ROT_QUAT || rot > RDB_ROT_RP ||
        (frame != 0 && frame != 1) || !(mass > 0.0))
        return RDB_ERR_ARG;
    *model = nullptr;
    const int n = 9 + (rot == RDB_ROT_QUAT ? 4 : 3);
    if (custom_check(n, m, wrench_body, np, RDB_F64, rot, frame) != 0) return RDB_ERR_COMPILE;
    rdb_model M;
    std::memset(&M, 0, sizeof(M));
    M.ctx = ctx; M.kind = RDB_CUSTOM; M.rot = rot; M.frame = frame; M.n = n; M.m = m; M.nerr = 12;
    M.p.mass = mass; M.p.inv_mass = 1.0 / mass;
    for (int i = 0; i < 9; ++i) M.p.J[i] = J[i];
    inv3(M.p.J, M.p.Jinv);
    M.lie = lie_parts_rigid();
    M.custom = custom_create(n, m, wrench_body, params, np, rot, frame, &M.p);
    if (!M.custom) return RDB_ERR_ARG;
    *model = new (std::nothrow) rdb_model(M);
    return *model ? 0 : RDB_ERR_ARG;
}

int rdb_custom_check(int n, int m, const char* f_body, int nparams, int dtype) {
    return custom_check(n, m, f_body, nparams, dtype) == 0 ? 0 : RDB_ERR_COMPILE;
}

int rdb_custom_rigid_check(int rot, int frame, int m, const char* wrench_body, int nparams, int dtype) {
    if (rot < RDB_ROT_QUAT || rot > RDB_ROT_RP || (frame != 0 && frame != 1) || m < 1) return RDB_ERR_ARG;
    return custom_check(9 + (rot == RDB_ROT_QUAT ? 4 : 3), m, wrench_body, nparams, dtype, rot, frame) == 0 ? 0 : RDB_ERR_COMPILE;
}

const char* rdb_last_log(void) { return custom_last_log(); }

int rdb_model_destroy(rdb_model* m) {
    if (m && m->custom) custom_destroy(m->custom);
    delete m;
    return 0;
}

int rdb_model_dims(const rdb_model* M, int* n, int* m, int* nerr) {
    if (!M) return RDB_ERR_ARG;
    if (n) *n = M->n;
    if (m) *m = M->m;
    if (nerr) *nerr = M->nerr;
    return 0;
}

int rdb_dynamics(const rdb_model* M, int dtype, int layout, int64_t N, const void* Z, const double* t, void* xdot, void* stream) {
    return knot_op(M, Q_CONTINUOUS, dtype, layout, 0, N, Z, t, nullptr, 0.0, nullptr, xdot, stream);
}

int rdb_discrete_dynamics(const rdb_model* M, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* t,
                          const double* dt, double dt0, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (Q < 0) return RDB_ERR_ARG;
    return knot_op(M, Q, dtype, layout, 0, N, Z, t, dt, dt0, nullptr, xn, stream);
}

int rdb_jacobian(const rdb_model* M, int dtype, int layout, int64_t N, const void* Z, const double* t, void* J, void* xdot, void* stream) {
    return knot_op(M, Q_CONTINUOUS, dtype, layout, 1, N, Z, t, nullptr, 0.0, J, xdot, stream);
}

int rdb_discrete_jacobian(const rdb_model* M, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* t,
                          const double* dt, double dt0, void* J, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (Q < 0) return RDB_ERR_ARG;
    return knot_op(M, Q, dtype, layout, 1, N, Z, t, dt, dt0, J, xn, stream);
}

int rdb_discrete_error_jacobian(const rdb_model* M, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* t,
                                const double* dt, double dt0, void* Jbar, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (Q < 0) return RDB_ERR_ARG;
    return knot_op(M, Q, dtype, layout, 1, N, Z, t, dt, dt0, Jbar, xn, stream, 1);
}

int rdb_errstate_jacobian(const rdb_model* M, int dtype, int64_t N, const void* X, int ldx, void* G, void* stream) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ldx < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!X || !G) return RDB_ERR_ARG;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    const int kind = classify({X, G}, c->device);
    if (kind < 0) return kind;
    if (kind == 2) return lie_errstate_jacobian(dtype, M->rot, M->lie, M->n, M->nerr, N, X, ldx, G, c->sm_count, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    const void* dX = st.in(B_Z, X, size_t(N) * ldx * es);
    void* dG = st.outbuf(B_J, size_t(N) * M->n * M->nerr * es);
    if (!st.rc) st.rc = lie_errstate_jacobian(dtype, M->rot, M->lie, M->n, M->nerr, N, dX, ldx, dG, c->sm_count, st.s.st);
    st.back(G, B_J, size_t(N) * M->n * M->nerr * es);
    return st.finish();
}

int rdb_grad_errstate_jacobian(const rdb_model* M, int dtype, int64_t N, const void* X, int ldx, const void* Xbar, int ldb,
                               void* H, void* stream) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ldx < M->n || ldb < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!X || !Xbar || !H) return RDB_ERR_ARG;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    const int kind = classify({X, Xbar, H}, c->device);
    if (kind < 0) return kind;
    if (kind == 2) return lie_grad_errstate_jacobian(dtype, M->rot, M->lie, M->n, M->nerr, N, X, ldx, Xbar, ldb, H, c->sm_count, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    const void* dX = st.in(B_Z, X, size_t(N) * ldx * es);
    const void* dB = st.in(B_AUX, Xbar, size_t(N) * ldb * es);
    void* dH = st.outbuf(B_J, size_t(N) * M->nerr * M->nerr * es);
    if (!st.rc) st.rc = lie_grad_errstate_jacobian(dtype, M->rot, M->lie, M->n, M->nerr, N, dX, ldx, dB, ldb, dH, c->sm_count, st.s.st);
    st.back(H, B_J, size_t(N) * M->nerr * M->nerr * es);
    return st.finish();
}

int rdb_state_diff(const rdb_model* M, int dtype, int64_t N, const void* X, int ldx, const void* X0, int ldx0, void* dXo, void* stream) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ldx < M->n || ldx0 < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!X || !X0 || !dXo) return RDB_ERR_ARG;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    const int kind = classify({X, X0, dXo}, c->device);
    if (kind < 0) return kind;
    if (kind == 2) return lie_state_diff(dtype, M->rot, M->lie, M->n, M->nerr, N, X, ldx, X0, ldx0, dXo, c->sm_count, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    const void* dX = st.in(B_Z, X, size_t(N) * ldx * es);
    const void* dX0 = st.in(B_AUX, X0, size_t(N) * ldx0 * es);
    void* dD = st.outbuf(B_OUT, size_t(N) * M->nerr * es);
    if (!st.rc) st.rc = lie_state_diff(dtype, M->rot, M->lie, M->n, M->nerr, N, dX, ldx, dX0, ldx0, dD, c->sm_count, st.s.st);
    st.back(dXo, B_OUT, size_t(N) * M->nerr * es);
    return st.finish();
}

int rdb_rollout(const rdb_model* M, int integrator, int dtype, int64_t ntraj, int K, const void* x0, const void* U, const double* t,
                const double* dt, double dt0, void* X, void* stream) {
    const int Q = map_q(integrator);
    if (!M || Q < 0 || ntraj < 0 || K < 1 || (dtype != RDB_F32 && dtype != RDB_F64)) return RDB_ERR_ARG;
    if (ntraj == 0) return 0;
    if (!x0 || !X || (K > 1 && !U)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    if (!model_uses_time(M)) t = nullptr;
    const int kind = classify({x0, U, t, dt, X}, c->device);
    if (kind < 0) return kind;
    KnotRequest r;
    std::memset(&r, 0, sizeof(r));
    r.op = OP_ROLLOUT; r.Q = Q; r.dtype = dtype; r.params = M->p; r.dt0 = dt0; r.ntraj = ntraj; r.K = K;
    r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
    if (kind == 2) {
        r.x0 = x0; r.U = U; r.dt = dt; r.t = t; r.X = X; r.stream = (cudaStream_t)stream;
        return dispatch(M, dtype, &r);
    }
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    r.x0 = st.in(B_Z, x0, size_t(ntraj) * M->n * es);
    r.U = st.in(B_AUX, U, size_t(ntraj) * (K - 1) * M->m * es);
    r.dt = (const double*)st.in(B_DT, dt, size_t(ntraj) * K * 8);
    r.t = (const double*)st.in(B_T, t, size_t(ntraj) * K * 8);
    r.X = st.outbuf(B_J, size_t(ntraj) * K * M->n * es);
    r.stream = st.s.st;
    if (!st.rc) st.rc = dispatch(M, dtype, &r);
    st.back(X, B_J, size_t(ntraj) * K * M->n * es);
    return st.finish();
}


// dynamics_error(dmodel, z2, z1) / dynamics_error_jacobian!(sig, diff, dmodel, J2, J1, y2, y1, z2, z1) for N pairs of knot points
static int dynamics_error_op(const rdb_model* M, int integrator, int dtype, int64_t N, const void* Z1, const void* Z2, int ld2, const double* t,
                             const double* dt, double dt0, void* J2, void* J1, void* e, void* stream, int with_j) {
    const int Q = map_q(integrator);
    if (!M || Q < 0 || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ld2 < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!Z1 || !Z2 || (with_j && !J1 && !J2) || (!with_j && !e)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    if (!model_uses_time(M)) t = nullptr;
    const int kind = classify({Z1, Z2, t, dt, J2, J1, e}, c->device);
    if (kind < 0) return kind;
    const size_t es = esize(dtype);
    const int n = M->n, NZ = M->n + M->m;
    auto run = [&](const void* dZ1, const void* dZ2, const double* dtt, const double* ddt, void* dJ2, void* dJ1, void* de, cudaStream_t st) -> int {
        if (Q == Q_IMPLICIT_MIDPOINT) {
            KnotRequest r;
            std::memset(&r, 0, sizeof(r));
            r.op = OP_DYNERR; r.Q = Q; r.dtype = dtype; r.with_j = with_j; r.params = M->p; r.dt0 = dt0; r.dt = ddt; r.t = dtt;
            r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
            r.Z = dZ1; r.Z2 = dZ2; r.ld2 = ld2; r.J2 = dJ2; r.J = dJ1; r.out = de; r.N = N; r.stream = st;
            return dispatch(M, dtype, &r);
        }
        // explicit rules: e = discrete_dynamics(z1) - x2, J1 = the discrete Jacobian, J2 = [-I 0]   (src/discrete_dynamics.jl:137-138,181-182)
        if (with_j && dJ1) { if (int rc = knot_op(M, Q, dtype, RDB_AOS, 1, N, dZ1, dtt, ddt, dt0, dJ1, de, st)) return rc; }
        else if (de) { if (int rc = knot_op(M, Q, dtype, RDB_AOS, 0, N, dZ1, dtt, ddt, dt0, nullptr, de, st)) return rc; }
        if (!de && !dJ2) return 0;
        const long long total = N * (long long)(dJ2 ? n * NZ : n);
        const unsigned g = unsigned(total + 255 < (1ll << 28) ? (total + 255) / 256 : (1ll << 20));
        if (dtype == RDB_F32) explicit_error_fixup_kernel<float><<<g, 256, 0, st>>>(n, M->m, N, (const float*)dZ2, ld2, (float*)de, (float*)dJ2);
        else explicit_error_fixup_kernel<double><<<g, 256, 0, st>>>(n, M->m, N, (const double*)dZ2, ld2, (double*)de, (double*)dJ2);
        return int(cudaGetLastError());
    };
    if (kind == 2) return run(Z1, Z2, t, dt, J2, J1, e, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    Staged st(c->slot[0]);
    Slot& s2 = c->slot[1];                               // second set of staging buffers for Z2 and J2
    const void* dZ1 = st.in(B_Z, Z1, size_t(N) * NZ * es);
    const double* ddt = (const double*)st.in(B_DT, dt, size_t(N) * 8);
    const double* dtt = (const double*)st.in(B_T, t, size_t(N) * 8);
    void* dJ1 = J1 ? st.outbuf(B_J, size_t(N) * n * NZ * es) : nullptr;
    void* de = e ? st.outbuf(B_OUT, size_t(N) * n * es) : nullptr;
    void* dZ2 = nullptr; void* dJ2 = nullptr;
    if (!st.rc) st.rc = ensure(s2, B_Z, size_t(N) * ld2 * es);
    if (!st.rc) { dZ2 = s2.buf[B_Z]; st.rc = cuda_rc(cudaMemcpyAsync(dZ2, Z2, size_t(N) * ld2 * es, cudaMemcpyHostToDevice, st.s.st)); }
    if (!st.rc && J2) { st.rc = ensure(s2, B_J, size_t(N) * n * NZ * es); dJ2 = s2.buf[B_J]; }
    if (!st.rc) st.rc = run(dZ1, dZ2, dtt, ddt, dJ2, dJ1, de, st.s.st);
    if (J1) st.back(J1, B_J, size_t(N) * n * NZ * es);
    if (e) st.back(e, B_OUT, size_t(N) * n * es);
    if (!st.rc && J2) st.rc = cuda_rc(cudaMemcpyAsync(J2, dJ2, size_t(N) * n * NZ * es, cudaMemcpyDeviceToHost, st.s.st));
    return st.finish();
}

int rdb_dynamics_error(const rdb_model* M, int integrator, int dtype, int64_t N, const void* Z1, const void* Z2, int ld2, const double* t,
                       const double* dt, double dt0, void* e, void* stream) {
    return dynamics_error_op(M, integrator, dtype, N, Z1, Z2, ld2, t, dt, dt0, nullptr, nullptr, e, stream, 0);
}
int rdb_dynamics_error_jacobian(const rdb_model* M, int integrator, int dtype, int64_t N, const void* Z1, const void* Z2, int ld2, const double* t,
                                const double* dt, double dt0, void* J2, void* J1, void* e, void* stream) {
    return dynamics_error_op(M, integrator, dtype, N, Z1, Z2, ld2, t, dt, dt0, J2, J1, e, stream, 1);
}

// ---- pre-validated launches ("plans") ---------------------------------------------------------------------------------------------
int rdb_plan_create(const rdb_model* M, int op, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* t,
                    const double* dt, double dt0, void* J, void* out, rdb_plan** plan) {
    if (!plan) return RDB_ERR_ARG;
    *plan = nullptr;
    if (!M || N <= 0 || (dtype != RDB_F32 && dtype != RDB_F64) || (layout != RDB_AOS && layout != RDB_SOA) || !Z) return RDB_ERR_ARG;
    int Q = Q_CONTINUOUS, with_j = 0, err = 0;
    switch (op) {
        case RDB_OP_DYNAMICS: break;
        case RDB_OP_JACOBIAN: with_j = 1; break;
        case RDB_OP_DISCRETE_DYNAMICS: Q = map_q(integrator); break;
        case RDB_OP_DISCRETE_JACOBIAN: Q = map_q(integrator); with_j = 1; break;
        case RDB_OP_DISCRETE_ERROR_JACOBIAN: Q = map_q(integrator); with_j = 1; err = (M->rot != RDB_ROT_NONE); break;
        default: return RDB_ERR_ARG;
    }
    if (Q < 0 || (with_j && !J) || (!with_j && !out)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    if (!model_uses_time(M)) t = nullptr;
    if (Q == Q_CONTINUOUS) { dt = nullptr; dt0 = 0.0; }
    const int kind = classify({Z, t, dt, J, out}, c->device);
    if (kind < 0) return kind;
    if (kind != 2) return RDB_ERR_POINTER_MIX;          // plans are for device-resident data
    rdb_plan* p = new (std::nothrow) rdb_plan();
    if (!p) return RDB_ERR_ARG;
    p->M = M; p->dtype = dtype; p->layout = layout;
    std::memset(&p->r, 0, sizeof(p->r));
    KnotRequest& r = p->r;
    r.op = OP_KNOT; r.Q = Q; r.dtype = dtype; r.with_j = with_j; r.err = err; r.params = M->p; r.dt0 = dt0;
    r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
    r.Z = Z; r.dt = dt; r.t = t; r.J = J; r.out = out; r.N = N;
    *plan = p;
    return 0;
}

int rdb_plan_launch(const rdb_plan* p, void* stream) {
    if (!p) return RDB_ERR_ARG;
    RDB_ON_DEVICE(p->M->ctx);
    KnotRequest r = p->r;
    r.stream = (cudaStream_t)stream;
    return p->layout == RDB_SOA ? dispatch_soa(p->M, p->dtype, r, r.N) : dispatch(p->M, p->dtype, &r);
}

int rdb_plan_set_shared(rdb_plan* p, int shared) {
    if (!p) return RDB_ERR_ARG;
    p->r.whole_sm = shared ? 1 : 0;
    return 0;
}

int rdb_plan_destroy(rdb_plan* p) { delete p; return 0; }

// ---- persistent device trajectory ------------------------------------------------------------------------------------------------------
int rdb_trajectory_create(const rdb_model* M, int dtype, int64_t ntraj, int K, rdb_trajectory** traj) {
    if (!traj) return RDB_ERR_ARG;
    *traj = nullptr;
    if (!M || ntraj < 1 || K < 1 || (dtype != RDB_F32 && dtype != RDB_F64)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    RDB_ON_DEVICE(M->ctx);
    rdb_trajectory* T = new (std::nothrow) rdb_trajectory();
    if (!T) return RDB_ERR_ARG;
    T->M = M; T->dtype = dtype; T->ntraj = ntraj; T->K = K;
    const size_t rows = size_t(ntraj) * K, zb = rows * (M->n + M->m) * esize(dtype);
    cudaError_t e = cudaMalloc(&T->Z, zb);
    if (e == cudaSuccess) e = cudaMalloc((void**)&T->t, rows * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&T->dt, rows * 8);
    if (e == cudaSuccess) e = cudaMemset(T->Z, 0, zb);
    if (e == cudaSuccess) e = cudaMemset(T->t, 0, rows * 8);
    if (e == cudaSuccess) e = cudaMemset(T->dt, 0, rows * 8);
    for (auto& st : T->aux) if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&T->ev_fork, cudaEventDisableTiming);
    for (auto& ev : T->ev_join) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { rdb_trajectory_destroy(T); return int(e); }
    *traj = T;
    return 0;
}

int rdb_trajectory_destroy(rdb_trajectory* T) {
    if (!T) return 0;
    DeviceGuard guard(T->M->ctx->device);
    for (auto& st : T->aux) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    if (T->ev_fork) cudaEventDestroy(T->ev_fork);
    for (auto& ev : T->ev_join) if (ev) cudaEventDestroy(ev);
    for (auto& ev : T->ev_chunk) if (ev) cudaEventDestroy(ev);
    if (T->Z) cudaFree(T->Z);
    if (T->t) cudaFree(T->t);
    if (T->dt) cudaFree(T->dt);
    if (T->stage) cudaFree(T->stage);
    delete T;
    return 0;
}

int rdb_trajectory_dims(const rdb_trajectory* T, int64_t* ntraj, int* K, int* n, int* m, int* dtype) {
    if (!T) return RDB_ERR_ARG;
    if (ntraj) *ntraj = T->ntraj;
    if (K) *K = T->K;
    if (n) *n = T->M->n;
    if (m) *m = T->M->m;
    if (dtype) *dtype = T->dtype;
    return 0;
}

int rdb_trajectory_data(const rdb_trajectory* T, void** Z, double** t, double** dt) {
    if (!T) return RDB_ERR_ARG;
    if (Z) *Z = T->Z;
    if (t) *t = T->t;
    if (dt) *dt = T->dt;
    return 0;
}

}  // extern "C"

namespace {
int traj_stage(rdb_trajectory* T, size_t bytes) {
    if (bytes <= T->stage_cap) return 0;
    if (T->stage) { RDB_CUDA(cudaFree(T->stage)); T->stage = nullptr; T->stage_cap = 0; }
    RDB_CUDA(cudaMalloc(&T->stage, bytes));
    T->stage_cap = bytes;
    return 0;
}
// copy a dense (rows, width) block `src` (host or device) into columns [col, col + width) of rows [r0, r0 + rows) of Z
int traj_put(rdb_trajectory* T, const void* src, long long r0, long long rows, int col, int width, cudaStream_t st) {
    if (!T || !src || rows < 0) return RDB_ERR_ARG;
    if (rows == 0) return 0;
    rdb_context* c = T->M->ctx;
    RDB_ON_DEVICE(c);
    const int kind = ptr_kind(src, c->device);
    if (kind < 0) return kind;
    const size_t es = esize(T->dtype);
    const int NZ = T->M->n + T->M->m;
    char* dst = (char*)T->Z + size_t(r0) * NZ * es;
    if (kind == 2) return copy_cols(T->dtype, src, width, 0, dst, NZ, col, width, rows, st);
    std::lock_guard<std::mutex> lock(T->mu);
    const size_t bytes = size_t(rows) * width * es;
    if (int rc = traj_stage(T, bytes)) return rc;
    RDB_CUDA(cudaMemcpyAsync(T->stage, src, bytes, cudaMemcpyHostToDevice, st));
    if (int rc = copy_cols(T->dtype, T->stage, width, 0, dst, NZ, col, width, rows, st)) return rc;
    return cuda_rc(cudaStreamSynchronize(st));           // host source: consumed when the call returns
}
int traj_get(rdb_trajectory* T, void* dstp, int col, int width, cudaStream_t st) {
    if (!T || !dstp) return RDB_ERR_ARG;
    rdb_context* c = T->M->ctx;
    RDB_ON_DEVICE(c);
    const int kind = ptr_kind(dstp, c->device);
    if (kind < 0) return kind;
    const int NZ = T->M->n + T->M->m;
    const long long rows = T->ntraj * T->K;
    if (kind == 2) return copy_cols(T->dtype, T->Z, NZ, col, dstp, width, 0, width, rows, st);
    std::lock_guard<std::mutex> lock(T->mu);
    const size_t bytes = size_t(rows) * width * esize(T->dtype);
    if (int rc = traj_stage(T, bytes)) return rc;
    if (int rc = copy_cols(T->dtype, T->Z, NZ, col, T->stage, width, 0, width, rows, st)) return rc;
    RDB_CUDA(cudaMemcpyAsync(dstp, T->stage, bytes, cudaMemcpyDeviceToHost, st));
    return cuda_rc(cudaStreamSynchronize(st));
}
KnotRequest traj_request(const rdb_trajectory* T, int Q, int err) {
    const rdb_model* M = T->M;
    rdb_context* c = M->ctx;
    KnotRequest r;
    std::memset(&r, 0, sizeof(r));
    r.Q = Q; r.dtype = T->dtype; r.params = M->p; r.dt0 = 0.0; r.dt = T->dt; r.t = model_uses_time(M) ? T->t : nullptr;
    r.err = (M->rot == RDB_ROT_NONE) ? 0 : err;
    r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
    return r;
}
}  // namespace

extern "C" {

int rdb_trajectory_set_states(rdb_trajectory* T, const void* X, void* stream) {
    return T ? traj_put(T, X, 0, T->ntraj * T->K, 0, T->M->n, (cudaStream_t)stream) : RDB_ERR_ARG;
}
int rdb_trajectory_set_initial_state(rdb_trajectory* T, const void* x0, void* stream) {
    return T ? traj_put(T, x0, 0, T->ntraj, 0, T->M->n, (cudaStream_t)stream) : RDB_ERR_ARG;
}
int rdb_trajectory_set_controls(rdb_trajectory* T, const void* U, int knots, void* stream) {
    if (!T || (knots != T->K && knots != T->K - 1)) return RDB_ERR_ARG;
    if (int rc = traj_put(T, U, 0, T->ntraj * knots, T->M->n, T->M->m, (cudaStream_t)stream)) return rc;
    if (knots == T->K - 1) {                             // no terminal control given: it is zero (src/knotpoint.jl:57-67)
        RDB_ON_DEVICE(T->M->ctx);
        return zero_cols(T->dtype, T->Z, T->M->n + T->M->m, T->M->n, T->M->m, T->ntraj * (T->K - 1), T->ntraj * T->K, (cudaStream_t)stream);
    }
    return 0;
}
int rdb_trajectory_set_timesteps(rdb_trajectory* T, const double* dt, double dt0, double t0, void* stream) {
    if (!T) return RDB_ERR_ARG;
    rdb_context* c = T->M->ctx;
    RDB_ON_DEVICE(c);
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = T->ntraj * T->K;
    const int kind = ptr_kind(dt, c->device);
    if (kind < 0) return kind;
    if (kind != 1) return time_grid(dt, dt0, t0, T->dt, T->t, T->ntraj, T->K, st);
    std::lock_guard<std::mutex> lock(T->mu);
    if (int rc = traj_stage(T, size_t(rows) * 8)) return rc;
    RDB_CUDA(cudaMemcpyAsync(T->stage, dt, size_t(rows) * 8, cudaMemcpyHostToDevice, st));
    if (int rc = time_grid((const double*)T->stage, dt0, t0, T->dt, T->t, T->ntraj, T->K, st)) return rc;
    return cuda_rc(cudaStreamSynchronize(st));
}
int rdb_trajectory_get_states(rdb_trajectory* T, void* X, void* stream) { return T ? traj_get(T, X, 0, T->M->n, (cudaStream_t)stream) : RDB_ERR_ARG; }
int rdb_trajectory_get_controls(rdb_trajectory* T, void* U, void* stream) { return T ? traj_get(T, U, T->M->n, T->M->m, (cudaStream_t)stream) : RDB_ERR_ARG; }

int rdb_trajectory_rollout(rdb_trajectory* T, int integrator, void* stream) {
    const int Q = map_q(integrator);
    if (!T || Q < 0) return RDB_ERR_ARG;
    RDB_ON_DEVICE(T->M->ctx);
    KnotRequest r = traj_request(T, Q, 0);
    r.op = OP_ROLLOUT; r.zmode = 1; r.kb = 0; r.ke = T->K - 1; r.X = T->Z; r.ntraj = T->ntraj; r.K = T->K; r.stream = (cudaStream_t)stream;
    return dispatch(T->M, T->dtype, &r);
}

int rdb_trajectory_linearize(rdb_trajectory* T, int integrator, int error_state, void* J, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (!T || Q < 0 || !J) return RDB_ERR_ARG;
    return knot_op(T->M, Q, T->dtype, RDB_AOS, 1, T->ntraj * T->K, T->Z, T->t, T->dt, 0.0, J, xn, stream, error_state ? 1 : 0);
}

// Forward pass + linearisation in one call.  Default (chunks <= 1): TWO PHASES on the caller's stream — the rollout kernel (latency-bound:
// one thread per trajectory, sequential in k) writes z = [x;u] rows in place, then ONE Jacobian launch streams all K * ntraj knots at
// full-GPU rate.  chunks > 1 is the measured alternative, a two-stream pipeline: the rollout runs in chunks of knots on one stream and the
// Jacobians of a finished chunk (a contiguous row range of the knot-major batch) on another while the next chunk is rolled out, forked
// from / joined into `stream` by events.  On B200 it LOSES (4096 x 256 quadrotor fp32: two-phase 0.44 ms, pipelined 0.61-0.77 ms;
// profiles/rollout_r02.md): a Jacobian CTA fills its SM's register file, so the two kernels exclude each other SM by SM, the rollout
// warps that do share an SM lose issue slots to FP32-issue-bound neighbours, and the sequential stage — the critical path — gets
// slower; packing the rollout onto 16 SMs and capping the Jacobian grid to the other 132 was worse still (0.67-0.70 ms).
int rdb_trajectory_rollout_linearize(rdb_trajectory* T, int integrator, int error_state, int chunks, void* J, void* stream) {
    const int Q = map_q(integrator);
    if (!T || Q < 0 || !J || chunks < 0) return RDB_ERR_ARG;
    const rdb_model* M = T->M;
    rdb_context* c = M->ctx;
    RDB_ON_DEVICE(c);
    const int jkind = ptr_kind(J, c->device);
    if (jkind < 0) return jkind;
    if (jkind != 2) {                                    // host Jacobians: roll out, then the chunked device -> host pipeline of knot_op
        if (int rc = rdb_trajectory_rollout(T, integrator, stream)) return rc;
        return rdb_trajectory_linearize(T, integrator, error_state, J, nullptr, stream);
    }
    if (chunks <= 1) {
        if (int rc = rdb_trajectory_rollout(T, integrator, stream)) return rc;
        return rdb_trajectory_linearize(T, integrator, error_state, J, nullptr, stream);
    }
    std::lock_guard<std::mutex> lock(T->mu);
    const int steps = T->K - 1;
    int nch = chunks;
    if (nch > steps) nch = steps > 0 ? steps : 1;
    while ((int)T->ev_chunk.size() < nch) {
        cudaEvent_t ev;
        RDB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        T->ev_chunk.push_back(ev);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int err = (M->rot == RDB_ROT_NONE) ? 0 : (error_state ? 1 : 0);
    const size_t es = esize(T->dtype);
    const int NZ = M->n + M->m, E = err ? M->nerr * (M->nerr + M->m) : M->n * NZ;
    RDB_CUDA(cudaEventRecord(T->ev_fork, st));
    for (auto& a : T->aux) RDB_CUDA(cudaStreamWaitEvent(a, T->ev_fork, 0));
    long long row_lo = 0;                                // first knot whose Jacobian has not been enqueued yet
    for (int ch = 0; ch < nch; ++ch) {
        const int kb = int((long long)steps * ch / nch), ke = int((long long)steps * (ch + 1) / nch);
        if (ke > kb) {
            KnotRequest r = traj_request(T, Q, 0);
            r.op = OP_ROLLOUT; r.zmode = 1; r.kb = kb; r.ke = ke; r.X = T->Z; r.ntraj = T->ntraj; r.K = T->K; r.stream = T->aux[0];
            if (int rc = dispatch(M, T->dtype, &r)) return rc;
        }
        RDB_CUDA(cudaEventRecord(T->ev_chunk[ch], T->aux[0]));
        RDB_CUDA(cudaStreamWaitEvent(T->aux[1], T->ev_chunk[ch], 0));
        const long long row_hi = (ch == nch - 1) ? (long long)T->K : (long long)ke + 1;   // states of knots < row_hi are final
        if (row_hi > row_lo) {
            KnotRequest r = traj_request(T, Q, err);
            r.op = OP_KNOT; r.with_j = 1; r.N = (row_hi - row_lo) * T->ntraj; r.stream = T->aux[1];
            r.Z = (const char*)T->Z + size_t(row_lo) * T->ntraj * NZ * es;
            r.dt = T->dt + row_lo * T->ntraj; if (r.t) r.t = T->t + row_lo * T->ntraj;
            r.J = (char*)J + size_t(row_lo) * T->ntraj * E * es;
            if (int rc = dispatch(M, T->dtype, &r)) return rc;
            row_lo = row_hi;
        }
    }
    for (int i = 0; i < 2; ++i) {
        RDB_CUDA(cudaEventRecord(T->ev_join[i], T->aux[i]));
        RDB_CUDA(cudaStreamWaitEvent(st, T->ev_join[i], 0));
    }
    return 0;
}

}  // extern "C"

