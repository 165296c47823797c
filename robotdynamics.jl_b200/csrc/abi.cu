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

using namespace rdb;

namespace {
constexpr int NSLOT = 3;                 // pipeline depth of the host-pointer path
constexpr long long HOST_CHUNK_DEFAULT = 1 << 16;  // knot points per pipelined chunk (RDB200_HOST_CHUNK overrides, for experiments)
enum { B_Z = 0, B_DT = 1, B_J = 2, B_OUT = 3, B_AUX = 4, NBUF = 5 };

struct Slot {
    cudaStream_t st = nullptr;
    void* buf[NBUF] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[NBUF] = {0, 0, 0, 0, 0};
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
    std::mutex mu;   // the staging slots are shared by all host-pointer calls on this context
};

struct rdb_model {
    rdb_context* ctx;
    int kind, rot, frame, D;
    int n, m, nerr;
    ModelParams<double> p;
    CustomModel* custom;   // kind == RDB_CUSTOM: NVRTC-compiled user model (custom.cu)
};

namespace {

int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : int(e); }
#define RDB_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return int(e__); } while (0)

void inv3(const double* A, double* Ai) {
    const double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], h = A[7], i = A[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    Ai[0] = (e * i - f * h) * id; Ai[1] = (c * h - b * i) * id; Ai[2] = (b * f - c * e) * id;
    Ai[3] = (f * g - d * i) * id; Ai[4] = (a * i - c * g) * id; Ai[5] = (c * d - a * f) * id;
    Ai[6] = (d * h - e * g) * id; Ai[7] = (b * g - a * h) * id; Ai[8] = (a * e - b * d) * id;
}

// 0 = null, 1 = host (pinned, registered or pageable), 2 = device / managed
int ptr_kind(const void* p) {
    if (!p) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 1; }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? 2 : 1;
}
// classify a set of data pointers: returns 1 host, 2 device, 0 all-null, RDB_ERR_POINTER_MIX on a mix
int classify(std::initializer_list<const void*> ps) {
    int kind = 0;
    for (const void* p : ps) {
        const int k = ptr_kind(p);
        if (k == 0) continue;
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
    SoaScratch* sc;
    { std::lock_guard<std::mutex> lock(c->soa_mu); sc = &c->soa[r.stream]; }
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
        if ((rc = dispatch(M, dtype, &q))) return rc;
        if (r.J && (rc = aos_to_soa(dtype, sc->buf[1], (char*)r.J + size_t(k0) * es, ld, E, cnt, r.stream))) return rc;
        if (r.out && (rc = aos_to_soa(dtype, sc->buf[2], (char*)r.out + size_t(k0) * es, ld, n, cnt, r.stream))) return rc;
    }
    return 0;
}

// The one knot-point operation behind rdb_dynamics / rdb_discrete_dynamics / rdb_jacobian / rdb_discrete_jacobian.
int knot_op(const rdb_model* M, int Q, int dtype, int layout, int with_j, long long N, const void* Z, const double* dt,
            double dt0, void* J, void* out, void* stream, int err = 0) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || (layout != RDB_AOS && layout != RDB_SOA)) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!Z || (with_j && !J) || (!with_j && !out)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    RDB_CUDA(cudaSetDevice(c->device));
    const int kind = classify({Z, dt, J, out});
    if (kind < 0) return kind;

    KnotRequest r;
    std::memset(&r, 0, sizeof(r));
    if (M->rot == RDB_ROT_NONE) err = 0;     // EuclideanState: G = I (src/statevectortype.jl:149-155)
    r.op = OP_KNOT; r.Q = Q; r.dtype = dtype; r.with_j = with_j; r.err = err; r.params = M->p;
    r.dt0 = dt0; r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
    if (kind == 2) {
        r.Z = Z; r.dt = dt; r.J = J; r.out = out; r.N = N; r.stream = (cudaStream_t)stream;
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
    for (long long k0 = 0; k0 < N && !rc; k0 += HOST_CHUNK, ++ci) {
        const long long cnt = (N - k0 < HOST_CHUNK) ? (N - k0) : HOST_CHUNK;
        Slot& s = c->slot[ci % NSLOT];
        if ((rc = ensure(s, B_Z, size_t(cap_knots) * NZ * es))) break;
        if (dt && (rc = ensure(s, B_DT, size_t(cap_knots) * 8))) break;
        if (J && (rc = ensure(s, B_J, size_t(cap_knots) * E * es))) break;
        if (out && (rc = ensure(s, B_OUT, size_t(cap_knots) * n * es))) break;
        if ((rc = copy_chunk(s.buf[B_Z], Z, layout, es, NZ, N, k0, cnt, cudaMemcpyHostToDevice, s.st))) break;
        if (dt && (rc = cuda_rc(cudaMemcpyAsync(s.buf[B_DT], dt + k0, size_t(cnt) * 8, cudaMemcpyHostToDevice, s.st)))) break;
        r.Z = s.buf[B_Z]; r.dt = dt ? (const double*)s.buf[B_DT] : nullptr;
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
    RDB_CUDA(cudaSetDevice(device));
    rdb_context* c = new (std::nothrow) rdb_context();
    if (!c) return RDB_ERR_ARG;
    c->device = device;
    RDB_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    if (const char* e = std::getenv("RDB200_PDL")) c->pdl = (e[0] != '0');
    if (const char* e = std::getenv("RDB200_HOST_CHUNK")) { const long long v = std::atoll(e); if (v >= 1024) c->host_chunk = v; }
    for (auto& s : c->slot) RDB_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    *ctx = c;
    return 0;
}

int rdb_destroy(rdb_context* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
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
    M.custom = custom_create(n, m, f_body, params, np);
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

int rdb_dynamics(const rdb_model* M, int dtype, int layout, int64_t N, const void* Z, const double* /*t*/, void* xdot, void* stream) {
    return knot_op(M, Q_CONTINUOUS, dtype, layout, 0, N, Z, nullptr, 0.0, nullptr, xdot, stream);
}

int rdb_discrete_dynamics(const rdb_model* M, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* /*t*/,
                          const double* dt, double dt0, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (Q < 0) return RDB_ERR_ARG;
    return knot_op(M, Q, dtype, layout, 0, N, Z, dt, dt0, nullptr, xn, stream);
}

int rdb_jacobian(const rdb_model* M, int dtype, int layout, int64_t N, const void* Z, const double* /*t*/, void* J, void* xdot, void* stream) {
    return knot_op(M, Q_CONTINUOUS, dtype, layout, 1, N, Z, nullptr, 0.0, J, xdot, stream);
}

int rdb_discrete_jacobian(const rdb_model* M, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* /*t*/,
                          const double* dt, double dt0, void* J, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (Q < 0) return RDB_ERR_ARG;
    return knot_op(M, Q, dtype, layout, 1, N, Z, dt, dt0, J, xn, stream);
}

int rdb_discrete_error_jacobian(const rdb_model* M, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* /*t*/,
                                const double* dt, double dt0, void* Jbar, void* xn, void* stream) {
    const int Q = map_q(integrator);
    if (Q < 0) return RDB_ERR_ARG;
    return knot_op(M, Q, dtype, layout, 1, N, Z, dt, dt0, Jbar, xn, stream, 1);
}

int rdb_errstate_jacobian(const rdb_model* M, int dtype, int64_t N, const void* X, int ldx, void* G, void* stream) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ldx < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!X || !G) return RDB_ERR_ARG;
    rdb_context* c = M->ctx;
    RDB_CUDA(cudaSetDevice(c->device));
    const int kind = classify({X, G});
    if (kind < 0) return kind;
    if (kind == 2) return lie_errstate_jacobian(dtype, M->rot, M->n, M->nerr, N, X, ldx, G, c->sm_count, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    const void* dX = st.in(B_Z, X, size_t(N) * ldx * es);
    void* dG = st.outbuf(B_J, size_t(N) * M->n * M->nerr * es);
    if (!st.rc) st.rc = lie_errstate_jacobian(dtype, M->rot, M->n, M->nerr, N, dX, ldx, dG, c->sm_count, st.s.st);
    st.back(G, B_J, size_t(N) * M->n * M->nerr * es);
    return st.finish();
}

int rdb_grad_errstate_jacobian(const rdb_model* M, int dtype, int64_t N, const void* X, int ldx, const void* Xbar, int ldb,
                               void* H, void* stream) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ldx < M->n || ldb < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!X || !Xbar || !H) return RDB_ERR_ARG;
    rdb_context* c = M->ctx;
    RDB_CUDA(cudaSetDevice(c->device));
    const int kind = classify({X, Xbar, H});
    if (kind < 0) return kind;
    if (kind == 2) return lie_grad_errstate_jacobian(dtype, M->rot, M->n, M->nerr, N, X, ldx, Xbar, ldb, H, c->sm_count, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    const void* dX = st.in(B_Z, X, size_t(N) * ldx * es);
    const void* dB = st.in(B_AUX, Xbar, size_t(N) * ldb * es);
    void* dH = st.outbuf(B_J, size_t(N) * M->nerr * M->nerr * es);
    if (!st.rc) st.rc = lie_grad_errstate_jacobian(dtype, M->rot, M->n, M->nerr, N, dX, ldx, dB, ldb, dH, c->sm_count, st.s.st);
    st.back(H, B_J, size_t(N) * M->nerr * M->nerr * es);
    return st.finish();
}

int rdb_state_diff(const rdb_model* M, int dtype, int64_t N, const void* X, int ldx, const void* X0, int ldx0, void* dXo, void* stream) {
    if (!M || N < 0 || (dtype != RDB_F32 && dtype != RDB_F64) || ldx < M->n || ldx0 < M->n) return RDB_ERR_ARG;
    if (N == 0) return 0;
    if (!X || !X0 || !dXo) return RDB_ERR_ARG;
    rdb_context* c = M->ctx;
    RDB_CUDA(cudaSetDevice(c->device));
    const int kind = classify({X, X0, dXo});
    if (kind < 0) return kind;
    if (kind == 2) return lie_state_diff(dtype, M->rot, M->n, M->nerr, N, X, ldx, X0, ldx0, dXo, c->sm_count, (cudaStream_t)stream);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    const void* dX = st.in(B_Z, X, size_t(N) * ldx * es);
    const void* dX0 = st.in(B_AUX, X0, size_t(N) * ldx0 * es);
    void* dD = st.outbuf(B_OUT, size_t(N) * M->nerr * es);
    if (!st.rc) st.rc = lie_state_diff(dtype, M->rot, M->n, M->nerr, N, dX, ldx, dX0, ldx0, dD, c->sm_count, st.s.st);
    st.back(dXo, B_OUT, size_t(N) * M->nerr * es);
    return st.finish();
}

int rdb_rollout(const rdb_model* M, int integrator, int dtype, int64_t ntraj, int K, const void* x0, const void* U, const double* /*t*/,
                const double* dt, double dt0, void* X, void* stream) {
    const int Q = map_q(integrator);
    if (!M || Q < 0 || ntraj < 0 || K < 1 || (dtype != RDB_F32 && dtype != RDB_F64)) return RDB_ERR_ARG;
    if (ntraj == 0) return 0;
    if (!x0 || !X || (K > 1 && !U)) return RDB_ERR_ARG;
    if (M->kind != RDB_CUSTOM && !find_unit(M->kind, M->rot, M->frame, M->D, dtype)) return RDB_ERR_NOT_IMPLEMENTED;
    rdb_context* c = M->ctx;
    RDB_CUDA(cudaSetDevice(c->device));
    const int kind = classify({x0, U, dt, X});
    if (kind < 0) return kind;
    KnotRequest r;
    std::memset(&r, 0, sizeof(r));
    r.op = OP_ROLLOUT; r.Q = Q; r.dtype = dtype; r.params = M->p; r.dt0 = dt0; r.ntraj = ntraj; r.K = K;
    r.dev = DeviceInfo{c->device, c->sm_count, c->pdl};
    if (kind == 2) {
        r.x0 = x0; r.U = U; r.dt = dt; r.X = X; r.stream = (cudaStream_t)stream;
        return dispatch(M, dtype, &r);
    }
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t es = esize(dtype);
    Staged st(c->slot[0]);
    r.x0 = st.in(B_Z, x0, size_t(ntraj) * M->n * es);
    r.U = st.in(B_AUX, U, size_t(ntraj) * (K - 1) * M->m * es);
    r.dt = (const double*)st.in(B_DT, dt, size_t(ntraj) * K * 8);
    r.X = st.outbuf(B_J, size_t(ntraj) * K * M->n * es);
    r.stream = st.s.st;
    if (!st.rc) st.rc = dispatch(M, dtype, &r);
    st.back(X, B_J, size_t(ntraj) * K * M->n * es);
    return st.finish();
}

}  // extern "C"
