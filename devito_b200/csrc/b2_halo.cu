// NCCL point-to-point halo exchange for x-slab decomposition, one process per GPU.
//
// Replaces the reference's generated MPI routines `haloupdate0/sendrecv0/gather0/scatter0`
// (devito/mpi/routines.py:285-552; printed in examples/mpi/overview.ipynb:503-560). Because
// the decomposed dimension x is the slowest-varying one of the (t, x, y, z) row-major
// layout, a face of `width` yz-planes is one contiguous block: no pack/unpack kernels
// (`gather0/scatter0`) are needed and NCCL sends/receives straight from/to field memory.
// Only `radius` planes are exchanged (the reference ships the full `space_order`-wide halo,
// devito/types/dense.py:1256-1259). Neighbours at the physical boundary are skipped
// (MPI_PROC_NULL in the reference, routines.py:429-433).
#include "b2_halo.cuh"
#include "b2_ptx.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl(const char *path) {
    if (g_nccl.dl) return B2_OK;
    const char *cands[] = {path, "libnccl.so.2", "libnccl.so"};
    void *dl = nullptr;
    for (const char *c : cands) {
        if (!c || !*c) continue;
        dl = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (dl) break;
    }
    if (!dl) { b2::set_error("cannot dlopen NCCL (%s)", dlerror()); return B2_ERR_COMM; }
#define LOAD(field, sym)                                                              \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(dl, sym));          \
    if (!g_nccl.field) { b2::set_error("NCCL symbol %s missing", sym); return B2_ERR_COMM; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.dl = dl;
    return B2_OK;
}

#define B2_NCCL(call)                                                                   \
    do {                                                                                \
        ncclResult_t _r = (call);                                                       \
        if (_r != ncclSuccess) {                                                        \
            b2::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                          g_nccl.GetErrorString(_r));                                   \
            return B2_ERR_COMM;                                                         \
        }                                                                               \
    } while (0)

}  // namespace

namespace b2 {

int halo_enqueue(b2_halo_ctx *ctx, float *base, size_t plane_elems, int lo, int n, int width) {
    const int left = ctx->rank - 1, right = ctx->rank + 1;
    const bool has_l = left >= 0, has_r = right < ctx->nranks;
    if (!has_l && !has_r) return B2_OK;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const size_t cnt = plane_elems * (size_t)width;
    B2_NCCL(g_nccl.GroupStart());
    if (has_l) {
        B2_NCCL(g_nccl.Send(base + (size_t)lo * plane_elems, cnt, ncclFloat, left, comm, ctx->comm_stream));
        B2_NCCL(g_nccl.Recv(base + (size_t)(lo - width) * plane_elems, cnt, ncclFloat, left, comm,
                            ctx->comm_stream));
    }
    if (has_r) {
        B2_NCCL(g_nccl.Send(base + (size_t)(lo + n - width) * plane_elems, cnt, ncclFloat, right, comm,
                            ctx->comm_stream));
        B2_NCCL(g_nccl.Recv(base + (size_t)(lo + n) * plane_elems, cnt, ncclFloat, right, comm,
                            ctx->comm_stream));
    }
    B2_NCCL(g_nccl.GroupEnd());
    return B2_OK;
}

// ---- peer-memory path ----------------------------------------------------------------------------
// Flags live in plain device memory that the neighbours map through CUDA IPC. The writer orders its
// peer stores (earlier kernels / copies of the same stream, or its own stores) before the flag with
// a system-scope release; the reader acquires at system scope before touching the halo planes.
__global__ void k_flag_wait(const int *flags, int want_left, int want_right) {
    // one thread; spins until both neighbours have signalled the required step
    if (want_left >= 0) while (b2ptx::ld_acquire_sys(flags + 0) < want_left) __nanosleep(40);
    if (want_right >= 0) while (b2ptx::ld_acquire_sys(flags + 1) < want_right) __nanosleep(40);
}

__global__ void k_flag_signal(int *left_remote, int *right_remote, int value) {
    // everything this stream did before (sweep kernel with its peer stores, injection, copies) is
    // ordered before the flag stores
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    if (left_remote) b2ptx::st_release_sys(left_remote, value);
    if (right_remote) b2ptx::st_release_sys(right_remote, value);
}

// After the boundary planes of time slot `slot1` are final (stencil strips + injection), store them
// into the neighbours' halos and signal. `base` = local field base (all slots).
static int p2p_push(b2_halo_ctx *ctx, const b2_halo_ctx::Reg &rg, const float *base, size_t slot_elems,
                    int slot1, size_t plane, int lo, int n, int width) {
    cudaStream_t st = ctx->p2p_async ? ctx->push_stream : stream();
    const size_t bytes = plane * (size_t)width * sizeof(float);
    const float *mine = base + (size_t)slot1 * slot_elems;
    // a neighbour's time slot holds (its owned planes + 2*halo) planes; halo width == lo here
    const size_t slot_l = (size_t)(rg.n_left + 2 * lo) * plane, slot_r = (size_t)(rg.n_right + 2 * lo) * plane;
    (void)slot_elems;
    if (rg.left) {      // my first `width` owned planes -> left neighbour's right halo
        float *dst = (float *)rg.left + (size_t)slot1 * slot_l + (size_t)(lo + rg.n_left) * plane;
        B2_CUDA(cudaMemcpyAsync(dst, mine + (size_t)lo * plane, bytes, cudaMemcpyDeviceToDevice, st), B2_ERR_COMM);
    }
    if (rg.right) {     // my last `width` owned planes -> right neighbour's left halo
        float *dst = (float *)rg.right + (size_t)slot1 * slot_r + (size_t)(lo - width) * plane;
        B2_CUDA(cudaMemcpyAsync(dst, mine + (size_t)(lo + n - width) * plane, bytes, cudaMemcpyDeviceToDevice, st),
                B2_ERR_COMM);
    }
    return B2_OK;
}

static int p2p_signal(b2_halo_ctx *ctx) {
    ++ctx->step;
    cudaStream_t st = ctx->p2p_async ? ctx->push_stream : stream();
    k_flag_signal<<<1, 1, 0, st>>>(ctx->flag_left_remote, ctx->flag_right_remote, ctx->step);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

static int p2p_wait(b2_halo_ctx *ctx) {
    const int want = ctx->step;       // neighbours run the same number of steps
    k_flag_wait<<<1, 1, 0, stream()>>>(ctx->flags_local, ctx->flag_left_remote ? want : -1,
                                        ctx->flag_right_remote ? want : -1);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

int halo_p2p_wait(b2_halo_ctx *ctx) { return p2p_wait(ctx); }

int halo_p2p_active(b2_halo_ctx *ctx, const void *base) { return ctx && ctx->p2p && ctx->find(base) != nullptr; }

// push the just-finished boundary planes of slot `slot1` of one or two fields, then signal once
int halo_p2p_publish(b2_halo_ctx *ctx, const float *f0, const float *f1, size_t slot_elems, int slot1,
                     size_t plane, int lo, int n, int width) {
    int rc;
    const float *fs[2] = {f0, f1};
    if (ctx->p2p_async) {
        // the planes are final on the main stream: hand them to the side stream
        B2_CUDA(cudaEventRecord(ctx->ev_final, stream()), B2_ERR_COMM);
        B2_CUDA(cudaStreamWaitEvent(ctx->push_stream, ctx->ev_final, 0), B2_ERR_COMM);
    }
    for (int i = 0; i < 2; ++i) {
        if (!fs[i]) continue;
        b2_halo_ctx::Reg *rg = ctx->find(fs[i]);
        if (!rg) { set_error("p2p: field not registered"); return B2_ERR_COMM; }
        if ((rc = p2p_push(ctx, *rg, fs[i], slot_elems, slot1, plane, lo, n, width))) return rc;
        rg->primed = true;
    }
    if ((rc = p2p_signal(ctx))) return rc;
    if (ctx->p2p_async) {
        B2_CUDA(cudaEventRecord(ctx->ev_push, ctx->push_stream), B2_ERR_COMM);
        ctx->push_pending = true;
    }
    return B2_OK;
}

int halo_p2p_drain(b2_halo_ctx *ctx) {
    if (!ctx || !ctx->push_pending) return B2_OK;
    // the pushed planes (time slot t1 of the previous step) are written again three steps later and
    // the library call must not return before its stores left: order the main stream after them
    B2_CUDA(cudaStreamWaitEvent(stream(), ctx->ev_push, 0), B2_ERR_COMM);
    ctx->push_pending = false;
    return B2_OK;
}

// planes a step needs from each neighbour: the stencil radius, twice that for the OT4 scheme (its update reads
// W = lap(u)/m at +-R, which reads u at +-2R)
int halo_width_iso(const IsoPlan &p) { return p.ot4 ? 2 * p.radius[0] : p.radius[0]; }

int halo_exchange_and_step_iso(b2_halo_ctx *ctx, const IsoPlan &p, int t0, int t2, int t1) {
    const int R = halo_width_iso(p);
    const int n = p.n[0];
    if (n < 2 * R) {
        set_error("halo: local slab of %d planes is thinner than twice the halo width %d", n, R);
        return B2_ERR_INVALID;
    }
    if (halo_p2p_active(ctx, p.u) && ctx->primed(p.u)) {
        // halos of u[t0] were stored by the neighbours at the end of their previous step: the
        // interior needs none of them; the boundary strips wait on the flags
        int rc;
        if ((rc = iso_step(p, t0, t2, t1, R, n - 2 * R))) return rc;
        if ((rc = halo_p2p_drain(ctx))) return rc;
        if ((rc = p2p_wait(ctx))) return rc;
        if ((rc = iso_step(p, t0, t2, t1, 0, R))) return rc;
        if ((rc = iso_step(p, t0, t2, t1, n - R, R))) return rc;
        return B2_OK;
    }
    cudaStream_t main = stream();
    B2_CUDA(cudaEventRecord(ctx->ev_ready, main), B2_ERR_COMM);
    B2_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0), B2_ERR_COMM);
    float *base = p.u + (size_t)t0 * p.slot_elems;
    int rc = halo_enqueue(ctx, base, (size_t)p.sx, p.o[0], n, R);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream), B2_ERR_COMM);
    if ((rc = iso_step(p, t0, t2, t1, R, n - 2 * R))) return rc;
    B2_CUDA(cudaStreamWaitEvent(main, ctx->ev_comm, 0), B2_ERR_COMM);
    if ((rc = iso_step(p, t0, t2, t1, 0, R))) return rc;
    if ((rc = iso_step(p, t0, t2, t1, n - R, R))) return rc;
    return B2_OK;
}

// ---- fused path (isotropic TMA sweep): ONE launch per step ----------------------------------------
// The sweep kernel itself stores its boundary planes into the neighbours' halos and acquires the
// neighbours' flags right before it reads their planes (b2_iso.cu, IsoTK). What is left on the host
// side is the flag release after the step's injection (halo_fused_signal). The first step of a call
// has no flags to rely on: it exchanges u[t0] through NCCL like the fallback path.
bool halo_fused_ok(b2_halo_ctx *ctx, const IsoPlan &p) {
    const char *e = getenv("B2_HALO_FUSED");          // read per call: tests and bench.py toggle the data path
    const bool off = e && atoi(e) == 0;
    return !off && halo_p2p_active(ctx, p.u) && p.use_tma && p.n[0] >= 2 * p.radius[0];
}

int halo_fuse_desc(b2_halo_ctx *ctx, const IsoPlan &p, IsoFuse &f) {
    const b2_halo_ctx::Reg *rg = ctx->find(p.u);
    if (!rg) { set_error("fused halo: field not registered"); return B2_ERR_COMM; }
    const long long plane = p.sx;
    f.peer_lo = (float *)rg->left;
    f.peer_hi = (float *)rg->right;
    f.n_lo = rg->n_left;
    f.n_hi = rg->n_right;
    f.slot_lo = (long long)(rg->n_left + 2 * p.o[0]) * plane;     // halo width == p.o[0] (x_m == 0 is enforced)
    f.slot_hi = (long long)(rg->n_right + 2 * p.o[0]) * plane;
    f.flag_lo = ctx->flag_left_remote ? ctx->flags_local + 0 : nullptr;
    f.flag_hi = ctx->flag_right_remote ? ctx->flags_local + 1 : nullptr;
    f.want = -1;
    return B2_OK;
}

int halo_step_iso_fused(b2_halo_ctx *ctx, const IsoPlan &p, int t0, int t2, int t1) {
    IsoFuse f;
    int rc = halo_fuse_desc(ctx, p, f);
    if (rc) return rc;
    if (ctx->primed(p.u)) {
        f.want = ctx->step;               // the neighbours released `step` after publishing u[t0]
    } else {
        cudaStream_t main = stream();
        B2_CUDA(cudaEventRecord(ctx->ev_ready, main), B2_ERR_COMM);
        B2_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0), B2_ERR_COMM);
        if ((rc = halo_enqueue(ctx, p.u + (size_t)t0 * p.slot_elems, (size_t)p.sx, p.o[0], p.n[0], p.radius[0])))
            return rc;
        B2_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream), B2_ERR_COMM);
        B2_CUDA(cudaStreamWaitEvent(main, ctx->ev_comm, 0), B2_ERR_COMM);
    }
    return iso_step_fused(p, t0, t2, t1, f);
}

// Exchange the boundary planes of time slot `t0` through NCCL (blocking the main stream) and mark the field as
// primed: what the first step of a call does, as a separate step for the streamed loop.
int halo_exchange_initial(b2_halo_ctx *ctx, const IsoPlan &p, int t0) {
    int rc = halo_exchange_slot(ctx, p, t0);
    if (rc) return rc;
    ctx->set_primed(p.u);
    return B2_OK;
}

// Blocking (w.r.t. the main stream) NCCL exchange of one time slot's boundary planes.
int halo_exchange_slot(b2_halo_ctx *ctx, const IsoPlan &p, int t0) {
    cudaStream_t main = stream();
    B2_CUDA(cudaEventRecord(ctx->ev_ready, main), B2_ERR_COMM);
    B2_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0), B2_ERR_COMM);
    int rc = halo_enqueue(ctx, p.u + (size_t)t0 * p.slot_elems, (size_t)p.sx, p.o[0], p.n[0], p.radius[0]);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream), B2_ERR_COMM);
    B2_CUDA(cudaStreamWaitEvent(main, ctx->ev_comm, 0), B2_ERR_COMM);
    return B2_OK;
}

int halo_fused_signal(b2_halo_ctx *ctx, const void *field) {
    int rc = p2p_signal(ctx);
    if (rc) return rc;
    ctx->set_primed(field);
    return B2_OK;
}

int halo_exchange_and_step_tti(b2_halo_ctx *ctx, const TtiPlan &p, int t0, int t2, int t1) {
    // TTI footprint: Laplacian star of radius R (the rotated derivatives stay within R-1)
    const int R = p.R;
    const int n = p.n[0];
    if (n < 2 * R) {
        set_error("halo: local slab of %d planes is thinner than 2*radius=%d", n, 2 * R);
        return B2_ERR_INVALID;
    }
    if (halo_p2p_active(ctx, p.u) && halo_p2p_active(ctx, p.v) && ctx->primed(p.u) && ctx->primed(p.v)) {
        int rc;
        if ((rc = tti_step(p, t0, t2, t1, R, n - 2 * R))) return rc;
        if ((rc = halo_p2p_drain(ctx))) return rc;
        if ((rc = p2p_wait(ctx))) return rc;
        if ((rc = tti_step(p, t0, t2, t1, 0, R))) return rc;
        if ((rc = tti_step(p, t0, t2, t1, n - R, R))) return rc;
        return B2_OK;
    }
    cudaStream_t main = stream();
    B2_CUDA(cudaEventRecord(ctx->ev_ready, main), B2_ERR_COMM);
    B2_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0), B2_ERR_COMM);
    int rc;
    if ((rc = halo_enqueue(ctx, p.u + (size_t)t0 * p.slot_elems, (size_t)p.sx, p.o[0], n, R))) return rc;
    if ((rc = halo_enqueue(ctx, p.v + (size_t)t0 * p.slot_elems, (size_t)p.sx, p.o[0], n, R))) return rc;
    B2_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream), B2_ERR_COMM);
    if ((rc = tti_step(p, t0, t2, t1, R, n - 2 * R))) return rc;
    B2_CUDA(cudaStreamWaitEvent(main, ctx->ev_comm, 0), B2_ERR_COMM);
    if ((rc = tti_step(p, t0, t2, t1, 0, R))) return rc;
    if ((rc = tti_step(p, t0, t2, t1, n - R, R))) return rc;
    return B2_OK;
}

}  // namespace b2

extern "C" {

int b2_nccl_unique_id(const char *nccl_lib, char id_out[128]) {
    int rc = load_nccl(nccl_lib);
    if (rc) return rc;
    ncclUniqueId id;
    B2_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id_out, &id, 128);
    return B2_OK;
}

b2_halo_ctx *b2_halo_create(const char *nccl_lib, const char id_in[128], int rank, int nranks,
                            int deviceid) {
    if (load_nccl(nccl_lib)) return nullptr;
    if (b2::use_device(deviceid)) return nullptr;
    b2_halo_ctx *ctx = new b2_halo_ctx();
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->deviceid = deviceid;
    ncclUniqueId id;
    memcpy(&id, id_in, 128);
    ncclComm_t comm;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) {
        b2::set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r));
        delete ctx;
        return nullptr;
    }
    ctx->comm = comm;
    cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming);
    return ctx;
}

void b2_halo_destroy(b2_halo_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->deviceid);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    if (ctx->comm) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->push_stream) {
        cudaStreamSynchronize(ctx->push_stream);
        cudaEventDestroy(ctx->ev_final);
        cudaEventDestroy(ctx->ev_push);
        cudaStreamDestroy(ctx->push_stream);
    }
    delete ctx;
}

int b2_ipc_get_handle(void *devptr, char out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t h;
    B2_CUDA(cudaIpcGetMemHandle(&h, devptr), B2_ERR_COMM);
    memcpy(out, &h, 64);
    return B2_OK;
}

void *b2_ipc_open(const char handle[64]) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { b2::set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return nullptr; }
    return p;
}

int b2_ipc_close(void *mapped) {
    B2_CUDA(cudaIpcCloseMemHandle(mapped), B2_ERR_COMM);
    return B2_OK;
}

int b2_halo_p2p_setup(b2_halo_ctx *ctx, void *flags_local, void *fl, void *fr) {
    if (!ctx || !flags_local) { b2::set_error("b2_halo_p2p_setup: NULL"); return B2_ERR_INVALID; }
    ctx->flags_local = (int *)flags_local;
    ctx->flag_left_remote = (int *)fl;
    ctx->flag_right_remote = (int *)fr;
    ctx->p2p = true;
    ctx->step = 0;
    ctx->reset_primed();
    const char *as = getenv("B2_P2P_ASYNC");
    if (as && atoi(as) != 0 && !ctx->push_stream) {
        B2_CUDA(cudaSetDevice(ctx->deviceid), B2_ERR_DEVICE);
        int lo_prio = 0, hi_prio = 0;
        B2_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio), B2_ERR_COMM);
        B2_CUDA(cudaStreamCreateWithPriority(&ctx->push_stream, cudaStreamNonBlocking, hi_prio), B2_ERR_COMM);
        B2_CUDA(cudaEventCreateWithFlags(&ctx->ev_final, cudaEventDisableTiming), B2_ERR_COMM);
        B2_CUDA(cudaEventCreateWithFlags(&ctx->ev_push, cudaEventDisableTiming), B2_ERR_COMM);
        ctx->p2p_async = true;
    }
    return B2_OK;
}

int b2_halo_p2p_register(b2_halo_ctx *ctx, void *local_base, void *left_base, void *right_base, int n_left,
                         int n_right) {
    if (!ctx || !ctx->p2p) { b2::set_error("b2_halo_p2p_register: p2p not set up"); return B2_ERR_INVALID; }
    for (auto &r : ctx->regs)
        if (r.local == local_base) { r = {local_base, left_base, right_base, n_left, n_right, false}; return B2_OK; }
    ctx->regs.push_back({local_base, left_base, right_base, n_left, n_right, false});
    return B2_OK;
}

int b2_halo_update(b2_halo_ctx *ctx, struct b2_dataobj *f, int slot, int width) {
    if (!ctx || !f || !f->dmap) { b2::set_error("b2_halo_update: needs a device-resident field"); return B2_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(ctx->deviceid), B2_ERR_DEVICE);
    // f is (t, x, y, z) or (x, y, z) when slot < 0
    const int nd = slot < 0 ? 3 : 4;
    const int *sz = f->size;
    const int ax = sz[nd - 3], ay = sz[nd - 2], az = sz[nd - 1];
    const size_t plane = (size_t)ay * az;
    const int so_l = f->hsize ? f->hsize[2 * (nd - 3)] : width;
    const int so_r = f->hsize ? f->hsize[2 * (nd - 3) + 1] : width;
    if (width > so_l || width > so_r) { b2::set_error("b2_halo_update: width %d exceeds halo", width); return B2_ERR_INVALID; }
    float *base = (float *)f->dmap + (slot < 0 ? 0 : (size_t)slot * ax * plane);
    cudaStream_t main = b2::stream();
    B2_CUDA(cudaEventRecord(ctx->ev_ready, main), B2_ERR_COMM);
    B2_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0), B2_ERR_COMM);
    int rc = b2::halo_enqueue(ctx, base, plane, so_l, ax - so_l - so_r, width);
    if (rc) return rc;
    B2_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream), B2_ERR_COMM);
    B2_CUDA(cudaStreamWaitEvent(main, ctx->ev_comm, 0), B2_ERR_COMM);
    B2_CUDA(cudaStreamSynchronize(main), B2_ERR_COMM);
    return B2_OK;
}

}  // extern "C"
