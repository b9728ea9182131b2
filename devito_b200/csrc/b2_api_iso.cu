// C-ABI entry point b2_iso_forward: the whole time loop of the reference's generated
// `Forward` function (examples/seismic/acoustic/operators.py:113-150 builds it; loop shape
// printed in examples/seismic/tutorials/08_snapshotting.ipynb:473-505):
//
//   for time in [time_m, time_M]:  t0 = time % T, t1 = (time+1) % T, t2 = (time+T-1) % T
//       [haloupdate u[t0]]            (devito/mpi/routines.py:457-510; here NCCL, b2_halo.cu)
//       section0: u[t1] = stencil(u[t0], u[t2], m, damp)
//       section1: u[t1] += inject(src[time])
//       section2: rec[time] = interpolate(u[t0] | u[t1])
#include "b2_iso.cuh"
#include "b2_sparse.cuh"
#include "b2_halo.cuh"
#include <vector>
#include <cstdlib>
#include <string>
#include <chrono>
#include <algorithm>

using namespace b2;

namespace b2 {

// NaN/Inf check (reference: errctl='max' checks every 100 steps, passes/iet/errors.py:59-85)
__global__ void k_check_finite(const float *__restrict__ f, size_t n, int *flag) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) bad |= !isfinite(f[i]);
    if (bad) atomicOr(flag, 1);
}

int check_finite(const float *f, size_t n, bool &bad) {
    static int *d_flag = nullptr;
    if (!d_flag) B2_CUDA(cudaMalloc(&d_flag, sizeof(int)), B2_ERR_MEMORY);
    B2_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), stream()), B2_ERR_MEMORY);
    k_check_finite<<<148 * 4, 256, 0, stream()>>>(f, n, d_flag);
    count_launch();
    int h = 0;
    B2_CUDA(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream()), B2_ERR_MEMORY);
    B2_CUDA(cudaStreamSynchronize(stream()), B2_ERR_DEVICE);
    bad = h != 0;
    return B2_OK;
}

// Imaging condition (reference: `Inc(grad, -u * v.dt2)`, acoustic/operators.py:222):
// grad[p] -= usave[time][p] * (v[t+1][p] - 2 v[t][p] + v[t-1][p]) / dt^2 over the iterated box.
struct ImgK {
    float *__restrict__ grad;
    const float *__restrict__ us;
    const float *__restrict__ v0;
    const float *__restrict__ v1;
    const float *__restrict__ v2;
    long long sx, sy, gsx, gsy;      // strides of the wavefields / of grad
    int n0, n1, n2, o0, o1, o2, g0, g1, g2;
    float inv_dt2;
};

__global__ void __launch_bounds__(256) k_imaging(ImgK k) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) {
        const long long i = (long long)(k.o0 + x) * k.sx + (long long)(k.o1 + y) * k.sy + (k.o2 + z);
        const long long g = (long long)(k.g0 + x) * k.gsx + (long long)(k.g1 + y) * k.gsy + (k.g2 + z);
        const float d2 = (k.v1[i] - 2.0f * k.v0[i] + k.v2[i]) * k.inv_dt2;
        k.grad[g] = k.grad[g] - k.us[i] * d2;
    }
}

// Section timing with a reusable pool of CUDA events: 4 events per time step.
struct StepEvents {
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    bool on = false;
    cudaEvent_t next() {
        if (used == ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        cudaEvent_t e = ev[used++];
        cudaEventRecord(e, stream());
        return e;
    }
};
static StepEvents g_step_events;

// Host-resident snapshots are not staged as a whole (nsnaps wavefields would not fit in HBM for a
// production run): each snapshot is written into one of two device buffers by k_snapshot and drained
// to the caller's array on a copy stream while the stencil keeps running — the out-of-core saving the
// reference does with asynchronous streaming (devito/passes/clusters/buffering.py, passes/iet/
// asynchrony.py). Only the iteration box travels (cudaMemcpy3DAsync), so the halo of the host
// snapshots keeps its values, exactly like `Eq(usave, u)` over the domain.
struct SnapStreamer {
    bool active = false;
    bool registered = false;         // we pinned the caller's array for the duration of the call
    void *host = nullptr;
    size_t host_bytes = 0;
    float *buf[2] = {nullptr, nullptr};
    size_t one = 0;                  // elements of one snapshot (with its halo)
    int count = 0;
    static cudaStream_t copy_stream;
    static cudaEvent_t filled[2], drained[2];

    int begin(void *host_base, size_t nbytes, size_t one_elems) {
        host = host_base;
        host_bytes = nbytes;
        one = one_elems;
        if (!copy_stream) {
            B2_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking), B2_ERR_DEVICE);
            for (int i = 0; i < 2; ++i) {
                B2_CUDA(cudaEventCreateWithFlags(&filled[i], cudaEventDisableTiming), B2_ERR_DEVICE);
                B2_CUDA(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming), B2_ERR_DEVICE);
            }
        }
        active = true;               // from here on end() has something to release
        count = 0;
        for (int i = 0; i < 2; ++i) B2_CUDA(cudaMalloc(&buf[i], one * sizeof(float)), B2_ERR_MEMORY);
        // pinned memory makes the copies truly asynchronous; pageable memory still works (the driver
        // stages it), already-pinned memory (e.g. torch pin_memory) is reported as such and left alone
        cudaError_t e = cudaHostRegister(host, host_bytes, cudaHostRegisterDefault);
        registered = (e == cudaSuccess);
        if (!registered) cudaGetLastError();
        return B2_OK;
    }

    // device buffer for the next snapshot, safe to overwrite on the compute stream
    int acquire(float **out) {
        const int b = count & 1;
        if (count >= 2) B2_CUDA(cudaStreamWaitEvent(stream(), drained[b], 0), B2_ERR_DEVICE);
        *out = buf[b];
        return B2_OK;
    }

    // the snapshot kernel has been enqueued: drain the box [d, d + n) of the buffer into host slot `index`
    int release(int index, const int *sz /* x,y,z extents of one snapshot */, const int *d, const int *n) {
        const int b = count & 1;
        B2_CUDA(cudaEventRecord(filled[b], stream()), B2_ERR_DEVICE);
        B2_CUDA(cudaStreamWaitEvent(copy_stream, filled[b], 0), B2_ERR_DEVICE);
        cudaMemcpy3DParms prm = {};
        const size_t pitch = (size_t)sz[2] * sizeof(float);
        prm.srcPtr = make_cudaPitchedPtr(buf[b], pitch, sz[2], sz[1]);
        prm.dstPtr = make_cudaPitchedPtr((float *)host + (size_t)index * one, pitch, sz[2], sz[1]);
        prm.srcPos = make_cudaPos((size_t)d[2] * sizeof(float), d[1], d[0]);
        prm.dstPos = prm.srcPos;
        prm.extent = make_cudaExtent((size_t)n[2] * sizeof(float), n[1], n[0]);
        prm.kind = cudaMemcpyDeviceToHost;
        B2_CUDA(cudaMemcpy3DAsync(&prm, copy_stream), B2_ERR_MEMORY);
        B2_CUDA(cudaEventRecord(drained[b], copy_stream), B2_ERR_DEVICE);
        ++count;
        return B2_OK;
    }

    int end() {
        if (!active) return B2_OK;
        active = false;
        cudaError_t e = cudaStreamSynchronize(copy_stream);
        for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; }
        if (registered) cudaHostUnregister(host);
        registered = false;
        if (e != cudaSuccess) { set_error("snapshot streaming: %s", cudaGetErrorString(e)); return B2_ERR_MEMORY; }
        return B2_OK;
    }
};
cudaStream_t SnapStreamer::copy_stream = nullptr;
cudaEvent_t SnapStreamer::filled[2] = {nullptr, nullptr};
cudaEvent_t SnapStreamer::drained[2] = {nullptr, nullptr};

}  // namespace b2


// ---------------------------------------------------------------------------------------------------------
// Streamed time loop for a HOST-STAGED apply (the reference's per-apply copy semantics,
// devito/passes/iet/definitions.py:636-671, without its serial copy-in / compute / copy-out):
//
// the x axis is cut into chunks of W planes. Chunk c of u (3 slots), damp and the parameter array travels
// host -> device on a copy stream while the compute stream already time-steps the chunks that arrived —
// skewed: in "phase" p, step s = 1..L updates x in [pW - (s-1)R, (p+1)W - (s-1)R), the dependency cone of
// the R-point stencil, so every read is of planes that are uploaded and at the right time level, and the 3
// rotating time slots are never overwritten early (the argument is spelled out in DESIGN.md §3.6). Planes
// left of (p+1)W - (L-1)R are final after phase p and travel device -> host on a second copy stream while
// the later phases compute (PCIe is full duplex). Injection deposits into the cells of the range just
// updated; receivers add the partial sum over the cells of that range to their trace.
//
// Whole-call time ~ compute + first upload chunk + last download chunk instead of upload + compute + download.
// ---------------------------------------------------------------------------------------------------------
namespace b2 {

static double g_last_profile[5] = {0, 0, 0, 0, 0};

struct StreamedLoop {
    cudaStream_t up = nullptr, down = nullptr;
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    int init() {
        if (!up) {
            B2_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking), B2_ERR_DEVICE);
            B2_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking), B2_ERR_DEVICE);
        }
        used = 0;
        return B2_OK;
    }
    cudaEvent_t next() {
        if (used == ev.size()) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            ev.push_back(e);
        }
        return ev[used++];
    }
};
static StreamedLoop g_streamed;

static bool streamed_wanted(const b2_iso_args *a, const IsoPlan &p, const DevArray &u, bool host_io) {
    const char *e = getenv("B2_STREAM");
    if (e && atoi(e) == 0) return false;
    const bool forced = e && atoi(e) == 2;                 // tests: also on small grids
    if (!(u.owned || host_io) || a->ndim != 3 || !p.use_tma || p.tsize != 3) return false;
    if (a->adjoint || a->free_surface || a->ot4 || a->born_U || a->snap || a->grad) return false;
    // under x-slab decomposition the sweep is skewed along y and every sub-launch does the fused halo step
    if (a->halo && !halo_fused_ok(a->halo, p)) return false;
    const int L = a->time_M - a->time_m + 1;
    const int dim = a->halo ? 1 : 0;
    if (L < 4 || p.n[dim] < 8 * p.radius[dim]) return false;
    return forced || u.nbytes >= ((size_t)1 << 30);
}

// `dim` = the axis the sweep is cut and skewed along: 0 (x; chunks are contiguous plane blocks) on a single device,
// 1 (y) under x-slab decomposition — orthogonal to the decomposed axis, so that every rank runs the SAME schedule
// and each sub-launch is an ordinary fused halo step (peer stores + flag acquire inside the sweep kernel) restricted
// to its rows. (Skewing along x would make a rank's low edge run ahead of its high edge in time, i.e. ahead of the
// neighbour it has to exchange with.) Chunks along y are strided: one 2-D copy per time slot and chunk, contiguous
// pieces of W rows x a2 floats.
static int iso_forward_streamed(const b2_iso_args *a, IsoPlan &p, FieldGeom g, DevArray &u, DevArray *damp,
                                DevArray *param, SparseDev &src, SparseDev &rec, float scalar_scale, float dt2,
                                bool damp_io, bool param_io) {
    int rc;
    StreamedLoop &S = g_streamed;
    if ((rc = S.init())) return rc;
    const bool decomposed = a->halo != nullptr;
    int dim = decomposed ? 1 : 0;
    if (!decomposed) if (const char *e = getenv("B2_STREAM_DIM")) dim = atoi(e) == 1 ? 1 : 0;
    const int n = p.n[dim], R = p.radius[dim], so = p.so, ad = p.a[dim], a0 = p.a[0];
    const int L = a->time_M - a->time_m + 1;
    const size_t plane = (size_t)p.sx, rowlen = (size_t)p.a[2];
    int W = 128;
    if (const char *e = getenv("B2_STREAM_W")) W = atoi(e);
    W = std::max(W, 4 * R);
    W = std::min(W, n);
    const int K = (n + W - 1) / W;                                   // upload chunks
    const int P = (n + (L - 1) * R + W - 1) / W;                     // phases of the skewed sweep
    cudaStream_t cs = stream();
    // x planes that travel (dim 1): the halo planes next to a neighbour rank hold ITS data, not the host's
    const int px0 = (dim == 1 && g.nb_lo) ? so : 0;
    const int px1 = (dim == 1 && g.nb_hi) ? so + p.n[0] : a0;

    // one chunk [lo, hi) (allocated index along `dim`) of a (slot, x, y, z) or (x, y, z) array, host <-> device
    auto copy_chunk = [&](float *dev, float *host, size_t slot_off, int lo, int hi, bool to_device, cudaStream_t st) -> int {
        if (hi <= lo) return B2_OK;
        if (dim == 0) {
            const size_t off = slot_off + (size_t)lo * plane, cnt = (size_t)(hi - lo) * plane * sizeof(float);
            B2_CUDA(cudaMemcpyAsync(to_device ? (void *)(dev + off) : (void *)(host + off),
                                    to_device ? (const void *)(host + off) : (const void *)(dev + off), cnt,
                                    to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st), B2_ERR_MEMORY);
        } else {
            const size_t off = slot_off + (size_t)px0 * plane + (size_t)lo * rowlen;
            B2_CUDA(cudaMemcpy2DAsync(to_device ? (void *)(dev + off) : (void *)(host + off), plane * sizeof(float),
                                      to_device ? (const void *)(host + off) : (const void *)(dev + off),
                                      plane * sizeof(float), (size_t)(hi - lo) * rowlen * sizeof(float),
                                      (size_t)(px1 - px0), to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st),
                    B2_ERR_MEMORY);
        }
        return B2_OK;
    };

    // everything the compute stream did so far (sparse tables, traces) precedes the copies
    cudaEvent_t ev_start = S.next();
    B2_CUDA(cudaEventRecord(ev_start, cs), B2_ERR_DEVICE);
    B2_CUDA(cudaStreamWaitEvent(S.up, ev_start, 0), B2_ERR_DEVICE);
    B2_CUDA(cudaStreamWaitEvent(S.down, ev_start, 0), B2_ERR_DEVICE);

    const int t0_first = ((a->time_m % 3) + 3) % 3;
    if (decomposed) {
        // the first sub-launch exchanges the initial level's boundary planes through NCCL (like the first step of a
        // resident call): they must be on the device, all rows, before anything else
        const int Rx = p.radius[0];
        const size_t so0 = (size_t)t0_first * p.slot_elems;
        const int lo_pl[2] = {so, so + p.n[0] - Rx};
        for (int i = 0; i < 2; ++i) {
            const size_t off = so0 + (size_t)lo_pl[i] * plane;
            B2_CUDA(cudaMemcpyAsync((float *)u.d + off, (const float *)u.h + off, (size_t)Rx * plane * sizeof(float),
                                    cudaMemcpyHostToDevice, S.up), B2_ERR_MEMORY);
        }
    }

    if (decomposed) {
        // ... and exchanged with the neighbours (NCCL, like the first step of a resident call) before the sweep starts
        cudaEvent_t ev_b = S.next();
        B2_CUDA(cudaEventRecord(ev_b, S.up), B2_ERR_DEVICE);
        B2_CUDA(cudaStreamWaitEvent(cs, ev_b, 0), B2_ERR_DEVICE);
        if ((rc = halo_exchange_initial(a->halo, p, t0_first))) return rc;
    }

    // ---- enqueue every upload now: the copy stream works through them in order ----
    std::vector<cudaEvent_t> up_ev(K);
    std::vector<int> ulo(K), uhi(K);
    for (int c = 0; c < K; ++c) {
        ulo[c] = c == 0 ? 0 : uhi[c - 1];
        uhi[c] = c == K - 1 ? ad : std::min(ad, so + std::min(n, (c + 1) * W) + R);
        for (int t = 0; t < 3; ++t)
            if ((rc = copy_chunk((float *)u.d, (float *)u.h, (size_t)t * p.slot_elems, ulo[c], uhi[c], true, S.up))) return rc;
        if (damp && damp_io && (rc = copy_chunk((float *)damp->d, (float *)damp->h, 0, ulo[c], uhi[c], true, S.up))) return rc;
        if (param && param_io && (rc = copy_chunk((float *)param->d, (float *)param->h, 0, ulo[c], uhi[c], true, S.up))) return rc;
        up_ev[c] = S.next();
        B2_CUDA(cudaEventRecord(up_ev[c], S.up), B2_ERR_DEVICE);
    }

    // traces are accumulated range by range: clear the rows this call writes
    if (rec.present) {
        const int r0 = std::max(a->time_m, 0), r1 = std::min(a->time_M, rec.nt - 1);
        if (r1 >= r0)
            B2_CUDA(cudaMemsetAsync((float *)rec.data.d + (size_t)r0 * rec.npoint_total, 0,
                                    (size_t)(r1 - r0 + 1) * rec.npoint_total * sizeof(float), cs), B2_ERR_MEMORY);
    }

    auto range_geom = [&](int ra, int rb) {
        FieldGeom q = g;
        q.lo[dim] = ra;
        q.hi[dim] = rb - 1;
        if (dim == 0) {
            q.nb_lo = ra > 0;        // an interior cut: the cells beyond it belong to another range
            q.nb_hi = rb < n;
            q.restrict_x = true;
        } else {
            q.cut_lo1 = ra > 0;
            q.cut_hi1 = rb < n;
            q.restrict_y = true;
        }
        return q;
    };
    IsoFuse fz;
    if (decomposed && (rc = halo_fuse_desc(a->halo, p, fz))) return rc;

    int dlo = 0;                                                     // allocated index (along dim) already sent back
    for (int ph = 0; ph < P; ++ph) {
        if (ph < K) {
            B2_CUDA(cudaStreamWaitEvent(cs, up_ev[ph], 0), B2_ERR_DEVICE);
            if (dim == 0) rc = iso_coef_tabulate_planes(p, ulo[ph], uhi[ph]);
            else rc = iso_coef_tabulate_rows(p, 0, a0, ulo[ph], uhi[ph]);
            if (rc) return rc;
            if (rec.present && !a->rec_toff) {
                // the initial time level is complete on the chunk that just arrived
                const int ra = std::max(0, ulo[ph] - so), rb = std::min(n, uhi[ph] - so);
                if (rb > ra)
                    if ((rc = launch_interp(rec, range_geom(ra, rb), p.u + (size_t)t0_first * p.slot_elems, nullptr, a->time_m)))
                        return rc;
            }
        }
        for (int s = 1; s <= L; ++s) {
            const int shift = (s - 1) * R;
            const int rb = std::min(n, (ph + 1) * W - shift);
            if (rb <= 0) break;
            const int ra = std::max(0, ph * W - shift);
            if (ra >= rb) continue;
            const int time = a->time_m + s - 1;
            const int t0 = ((time % 3) + 3) % 3, t1 = (((time + 1) % 3) + 3) % 3, t2 = (((time - 1) % 3) + 3) % 3;
            float *f1 = p.u + (size_t)t1 * p.slot_elems;
            FieldGeom q = range_geom(ra, rb);
            if (dim == 0) {
                if ((rc = iso_step(p, t0, t2, t1, ra, rb - ra))) return rc;
            } else {
                IsoPlan pr = p;                                   // the same plan restricted to the rows [ra, rb)
                pr.o[1] = p.o[1] + ra;
                pr.n[1] = rb - ra;
                if (decomposed) {
                    if ((rc = halo_step_iso_fused(a->halo, pr, t0, t2, t1))) return rc;
                    q.peer_lo = fz.peer_lo; q.peer_hi = fz.peer_hi;
                    q.off_lo = (long long)t1 * fz.slot_lo + (long long)(p.o[0] + fz.n_lo) * p.sx;
                    q.off_hi = (long long)t1 * fz.slot_hi + (long long)(p.o[0] - p.n[0]) * p.sx;
                    q.nown = p.n[0]; q.pw = p.radius[0];
                } else if ((rc = iso_step(pr, t0, t2, t1, 0, p.n[0]))) {
                    return rc;
                }
            }
            if ((rc = launch_inject(src, q, f1, nullptr, time, p.param_kind, p.param, scalar_scale, dt2))) return rc;
            if (decomposed) {
                if ((rc = halo_fused_signal(a->halo, p.u))) return rc;
                // receivers next to a slab boundary sample halo cells of the level just written: the neighbour's
                // stores of this very sub-launch must have landed
                if (rec.present && (rc = halo_p2p_wait(a->halo))) return rc;
            }
            // time level time+1 is now complete on [ra, rb)
            if (rec.present) {
                const int row = a->rec_toff ? time : time + 1;
                FieldGeom qi = q;
                qi.peer_lo = qi.peer_hi = nullptr;
                if (row <= a->time_M && (rc = launch_interp(rec, qi, f1, nullptr, row))) return rc;
            }
        }
        // what lies left of (ph+1)W - (L-1)R has seen all L steps: send it home while the sweep goes on
        int dhi = ph == P - 1 ? ad : std::min(ad, std::max(0, so + (ph + 1) * W - (L - 1) * R));
        if (ph == P - 1 || dhi - dlo >= W / 2) {
            if (dhi > dlo) {
                cudaEvent_t e = S.next();
                B2_CUDA(cudaEventRecord(e, cs), B2_ERR_DEVICE);
                B2_CUDA(cudaStreamWaitEvent(S.down, e, 0), B2_ERR_DEVICE);
                for (int t = 0; t < 3; ++t)
                    if ((rc = copy_chunk((float *)u.d, (float *)u.h, (size_t)t * p.slot_elems, dlo, dhi, false, S.down))) return rc;
                dlo = dhi;
            }
        }
    }
    // the call returns when the last planes are home
    cudaEvent_t ev_down = S.next();
    B2_CUDA(cudaEventRecord(ev_down, S.down), B2_ERR_DEVICE);
    B2_CUDA(cudaStreamWaitEvent(cs, ev_down, 0), B2_ERR_DEVICE);
    return B2_OK;
}

}  // namespace b2

extern "C" void b2_last_call_profile(double out[5]) {
    for (int i = 0; i < 5; ++i) out[i] = b2::g_last_profile[i];
}

extern "C" int b2_iso_forward(const struct b2_iso_args *a) {
    if (!a || !a->u) { set_error("b2_iso_forward: NULL args"); return B2_ERR_INVALID; }
    if (a->ndim != 2 && a->ndim != 3) { set_error("b2_iso_forward: ndim must be 2 or 3"); return B2_ERR_INVALID; }
    if (a->radius < 1 || a->radius > B2_MAX_RADIUS || a->radius > a->space_order) {
        set_error("b2_iso_forward: unsupported radius %d (space_order %d)", a->radius, a->space_order);
        return B2_ERR_INVALID;
    }
    if (a->time_M < a->time_m) return B2_OK;
    std::lock_guard<std::mutex> api_lock(api_mutex());
    if (int rc0 = use_device(a->deviceid)) return rc0;

    const int nd = a->ndim;
    const int so = a->space_order;
    int rc = B2_OK;

    DevArray u, damp, param, grad, usave, bornU, borndm, snap;
    bool staged_grad = false, staged_usave = false, staged_bornU = false, staged_borndm = false;
    bool staged_snap = false;
    SnapStreamer streamer;
    SparseDev src, rec;
    IsoPlan p;
    FieldGeom g;
    bool staged_u = false, staged_damp = false, staged_param = false;
    bool u_is_home = false;                       // the streamed loop already brought u back to the host
    bool u_host_io = false;                       // caller-provided device buffer next to the host array (host_io)
    auto call_t0 = std::chrono::steady_clock::now();
    for (double &v : g_last_profile) v = 0.0;

    auto cleanup = [&](int code) {
        // copy back what the reference would copy back (u and rec), free staged copies
        const auto d2h_t0 = std::chrono::steady_clock::now();
        int r1 = staged_u ? stage_out(u, (code == B2_OK || code == B2_ERR_NAN) && !u_is_home) : B2_OK;
        if (staged_u && u_host_io && !u_is_home && (code == B2_OK || code == B2_ERR_NAN)) {
            cudaError_t ce = cudaMemcpyAsync(u.h, u.d, u.nbytes, cudaMemcpyDeviceToHost, stream());
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream());
            if (ce != cudaSuccess) { set_error("b2_iso_forward: device -> host copy failed"); r1 = B2_ERR_MEMORY; }
        }
        g_last_profile[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - d2h_t0).count();
        g_last_profile[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - call_t0).count();
        if (staged_damp) stage_out(damp, false);
        if (staged_param) stage_out(param, false);
        if (staged_usave) stage_out(usave, false);
        int r3 = staged_grad ? stage_out(grad, code == B2_OK) : B2_OK;
        if (staged_snap) {
            const int r5 = stage_out(snap, code == B2_OK || code == B2_ERR_NAN);
            if (!r3) r3 = r5;
        }
        {
            const int r6 = streamer.end();
            if (!r3) r3 = r6;
        }
        if (staged_borndm) stage_out(borndm, false);
        if (staged_bornU) {
            const int r4 = stage_out(bornU, code == B2_OK || code == B2_ERR_NAN);
            if (!r3) r3 = r4;
        }
        if (code == B2_OK && r3) return r3;
        sparse_stage_out(src, false);
        int r2 = sparse_stage_out(rec, code == B2_OK || code == B2_ERR_NAN);
        if (code != B2_OK) return code;
        return r1 ? r1 : r2;
    };

    // A large host-staged wavefield is a candidate for the streamed time loop (copies overlapped with a
    // skewed sweep, iso_forward_streamed): its uploads are then issued chunk by chunk, not here. Whether the
    // call really streams is known once the plan exists; otherwise the deferred copies are issued below.
    // host_io: the caller passed BOTH a host array and its own device buffer for u (and possibly damp / the
    // parameter array): the call moves data -> dmap before and dmap -> data after the loop (streamed when
    // possible). Under x-slab decomposition this is how a host-staged apply keeps using the CUDA-IPC registered
    // device allocations of the peer-memory halo path.
    const bool host_io = a->host_io && a->u->dmap && a->u->data;
    const bool maybe_stream = [&] {
        const char *e = getenv("B2_STREAM");
        if (e && atoi(e) == 0) return false;
        if (nd != 3 || !a->u->data || !a->damp) return false;
        if (a->u->dmap && !host_io) return false;
        if (a->adjoint || a->free_surface || a->ot4 || a->born_U || a->born_dm || a->snap || a->grad || a->usave)
            return false;
        return true;
    }();
    call_t0 = std::chrono::steady_clock::now();
    if ((rc = stage_in(a->u, nd + 1, u, !maybe_stream))) return cleanup(rc);
    staged_u = true;
    u_host_io = host_io;
    if (a->damp) {
        if ((rc = stage_in(a->damp, nd, damp, !maybe_stream))) return cleanup(rc);
        staged_damp = true;
    }
    if (a->param_kind != B2_PARAM_SCALAR) {
        if (!a->param) { set_error("b2_iso_forward: param array missing"); return cleanup(B2_ERR_INVALID); }
        if ((rc = stage_in(a->param, nd, param, !maybe_stream))) return cleanup(rc);
        staged_param = true;
    }
    if ((rc = sparse_stage_in(a->src, nd, src, true))) return cleanup(rc);
    if ((rc = sparse_stage_in(a->rec, nd, rec, true))) return cleanup(rc);
    // the support of a sparse point reaches r cells beyond the iterated box: it must fit in the halo
    // (devito/operations/interpolators.py:28-37 `check_radius`)
    if ((src.present && src.r > so) || (rec.present && rec.r > so)) {
        set_error("b2_iso_forward: sparse radius %d exceeds the halo (space_order %d)",
                  src.present && src.r > so ? src.r : rec.r, so);
        return cleanup(B2_ERR_INVALID);
    }
    if (a->grad || a->usave) {
        if (!a->grad || !a->usave || nd != 3) {
            set_error("b2_iso_forward: the imaging condition needs both grad and usave (3-D)");
            return cleanup(B2_ERR_INVALID);
        }
        if ((rc = stage_in(a->grad, 3, grad, true))) return cleanup(rc);
        staged_grad = true;
        if ((rc = stage_in(a->usave, 4, usave, true))) return cleanup(rc);
        staged_usave = true;
    }

    const bool born = a->born_U || a->born_dm;
    if (born) {
        if (!a->born_U || !a->born_dm || nd != 3 || a->adjoint || a->grad || a->free_surface || a->ot4) {
            set_error("b2_iso_forward: Born modelling needs born_U and born_dm, 3-D, forward in time, and is "
                      "not combined with a free surface, OT4 or the imaging condition");
            return cleanup(B2_ERR_INVALID);
        }
        if ((rc = stage_in(a->born_U, 4, bornU, true))) return cleanup(rc);
        staged_bornU = true;
        if ((rc = stage_in(a->born_dm, 3, borndm, true))) return cleanup(rc);
        staged_borndm = true;
        for (int d = 0; d < 4; ++d)
            if (bornU.size[d] != u.size[d]) {
                set_error("b2_iso_forward: born_U must have the layout of u");
                return cleanup(B2_ERR_INVALID);
            }
    }

    int snap_h = 0;
    if (a->snap) {
        if (a->snap_factor < 1 || a->adjoint) {
            set_error("b2_iso_forward: snapshots need snap_factor >= 1 and forward time stepping");
            return cleanup(B2_ERR_INVALID);
        }
        const bool stream_out = !a->snap->dmap && a->snap->data && nd == 3 &&
                                !(getenv("B2_SNAP_STREAM") && atoi(getenv("B2_SNAP_STREAM")) == 0);
        if (stream_out) {
            // describe the array without allocating it on the device
            if (!a->snap->size) { set_error("b2_iso_forward: snapshot dataobj without `size`"); return cleanup(B2_ERR_INVALID); }
            snap.ndim = nd + 1;
            for (int d = 0; d <= nd; ++d) snap.size[d] = a->snap->size[d];
            snap.h = a->snap->data;
            snap.d = nullptr;
        } else {
            if ((rc = stage_in(a->snap, nd + 1, snap, true))) return cleanup(rc);
            staged_snap = true;
        }
        snap_h = a->snap->hsize ? a->snap->hsize[2] : (snap.size[1] - (u.size[1] - 2 * so)) / 2;
        for (int d = 0; d < nd; ++d)
            if (snap.size[d + 1] != u.size[d + 1] - 2 * so + 2 * snap_h) {
                set_error("b2_iso_forward: snapshot extent %d on dim %d does not match the grid", snap.size[d + 1], d);
                return cleanup(B2_ERR_INVALID);
            }
        if (a->time_m < 0 || a->time_M / a->snap_factor >= snap.size[0]) {
            set_error("b2_iso_forward: time_M=%d needs %d snapshots, the array holds %d", a->time_M,
                      a->time_M / a->snap_factor + 1, snap.size[0]);
            return cleanup(B2_ERR_INVALID);
        }
        if (stream_out) {
            const size_t one = (size_t)snap.size[1] * snap.size[2] * snap.size[3];
            if ((rc = streamer.begin(snap.h, one * snap.size[0] * sizeof(float), one))) return cleanup(rc);
        }
    }

    // ---- geometry in the internal 3-dim convention ----
    const int lo_in[3] = {a->x_m, a->y_m, a->z_m};
    const int hi_in[3] = {a->x_M, a->y_M, a->z_M};
    p.tsize = u.size[0];
    p.so = so;
    if (nd == 3) {
        for (int d = 0; d < 3; ++d) {
            p.a[d] = u.size[d + 1];
            p.n[d] = hi_in[d] - lo_in[d] + 1;
            p.o[d] = lo_in[d] + so;
            p.radius[d] = a->radius;
        }
    } else {
        p.a[0] = 1; p.n[0] = 1; p.o[0] = 0; p.radius[0] = 0;
        for (int d = 0; d < 2; ++d) {
            p.a[d + 1] = u.size[d + 1];
            p.n[d + 1] = hi_in[d] - lo_in[d] + 1;
            p.o[d + 1] = lo_in[d] + so;
            p.radius[d + 1] = a->radius;
        }
    }
    for (int d = 0; d < 3; ++d) {
        if (p.n[d] <= 0) return cleanup(B2_OK);
        const int rd = p.radius[d];
        if (p.o[d] - rd < 0 || p.o[d] + p.n[d] - 1 + rd >= p.a[d]) {
            set_error("b2_iso_forward: iteration range [%d,%d] + radius %d leaves the allocated "
                      "array (extent %d) on dim %d", lo_in[nd == 3 ? d : d - 1], hi_in[nd == 3 ? d : d - 1],
                      rd, p.a[d], d);
            return cleanup(B2_ERR_INVALID);
        }
    }
    p.sy = p.a[2];
    p.sx = (long long)p.a[1] * p.a[2];
    p.slot_elems = (size_t)p.a[0] * p.a[1] * p.a[2];
    p.u = (float *)u.d;
    p.damp = a->damp ? (const float *)damp.d : nullptr;
    p.param = a->param_kind != B2_PARAM_SCALAR ? (const float *)param.d : nullptr;
    p.param_kind = a->param_kind;
    p.vp = a->vp;
    p.dt = a->dt;
    for (int d = 0; d < nd; ++d) {
        const int di = nd == 3 ? d : d + 1;
        if (!a->w[d]) { set_error("b2_iso_forward: weights for dim %d missing", d); return cleanup(B2_ERR_INVALID); }
        for (int i = 0; i <= a->radius; ++i) p.w[di][i] = a->w[d][i];
    }
    p.ot4 = a->ot4 != 0;
    if (p.ot4 && (a->free_surface || a->grad)) {
        set_error("b2_iso_forward: OT4 is not combined with a free surface or the imaging condition in this version");
        return cleanup(B2_ERR_INVALID);
    }
    p.defer_coef = maybe_stream || host_io;       // the arrays are not on the device yet
    if ((rc = iso_plan_init(p, a->kernel))) return cleanup(rc);
    const bool damp_io = staged_damp && (damp.owned || (host_io && a->damp->dmap && a->damp->data));
    const bool param_io = staged_param && (param.owned || (host_io && a->param->dmap && a->param->data));
    const bool streamed = maybe_stream && streamed_wanted(a, p, u, host_io);
    if ((maybe_stream || host_io) && !streamed) {
        // not streaming after all: the classic order — whole arrays in, tabulate, loop, whole arrays out
        DevArray *arrs[3] = {&u, damp_io ? &damp : nullptr, param_io ? &param : nullptr};
        for (DevArray *x : arrs)
            if (x && (x->owned || host_io))
                if (cudaMemcpyAsync(x->d, x->h, x->nbytes, cudaMemcpyHostToDevice, stream()) != cudaSuccess) {
                    set_error("b2_iso_forward: host -> device copy failed");
                    return cleanup(B2_ERR_MEMORY);
                }
        p.defer_coef = false;
        if (p.use_tma && (rc = iso_coef_tabulate_planes(p, 0, p.a[0]))) return cleanup(rc);
    }
    // Born: a second plan for the linearised field (same geometry and coefficient tables, own tensor maps)
    IsoPlan pU = p;
    int dmh = 0;
    if (born) {
        pU.u = (float *)bornU.d;
        if ((rc = iso_plan_init(pU, a->kernel))) return cleanup(rc);
        dmh = a->born_dm->hsize ? a->born_dm->hsize[0] : (borndm.size[0] - (u.size[1] - 2 * so)) / 2;
        for (int d = 0; d < 3; ++d)
            if (borndm.size[d] != u.size[d + 1] - 2 * so + 2 * dmh) {
                set_error("b2_iso_forward: born_dm extent %d on dim %d does not match the grid", borndm.size[d], d);
                return cleanup(B2_ERR_INVALID);
            }
    }
    if (a->free_surface && p.o[2] != so) {
        set_error("b2_iso_forward: a free surface needs the vertical iteration to start at 0");
        return cleanup(B2_ERR_INVALID);
    }

    g.sx = p.sx;
    g.sy = p.sy;
    g.slot_elems = p.slot_elems;
    g.so = so;
    g.ndim = nd;
    for (int d = 0; d < nd; ++d) { g.lo[d] = lo_in[d]; g.hi[d] = hi_in[d]; }
    if (a->halo) { g.nb_lo = a->halo->rank > 0; g.nb_hi = a->halo->rank < a->halo->nranks - 1; }

    const float dt2 = a->dt * a->dt;
    const float scalar_scale = dt2 * a->vp * a->vp;
    const int T = p.tsize;
    const bool timing = a->timers != nullptr;
    StepEvents &se = g_step_events;
    se.used = 0;
    const int nsteps = a->time_M - a->time_m + 1;
    // The reference reports section0 (stencil), section1 (injection), section2 (interpolation)
    // (devito/operator/profiling.py:40-123). Four event records per step would cost ~1 % of a step, so the split is
    // SAMPLED: every 16th step is bracketed (every step with B2_PROFILING=advanced), the sparse sections are
    // scaled up, and section0 is the loop's total minus them.
    static const bool adv = getenv("B2_PROFILING") && std::string(getenv("B2_PROFILING")) == "advanced";
    const int sample_stride = adv ? 1 : 16;
    static cudaEvent_t ev_tot[2] = {nullptr, nullptr}, ev_prof[2] = {nullptr, nullptr};
    if (!ev_tot[0]) {
        for (auto &e : ev_tot) cudaEventCreate(&e);
        for (auto &e : ev_prof) cudaEventCreate(&e);
    }
    int step_index = 0, nsampled = 0;
    if (timing) cudaEventRecord(ev_tot[0], stream());

    const bool p2p = a->halo && halo_p2p_active(a->halo, p.u);
    // halo step fused into the sweep kernel (peer stores + flag acquire inside k_iso_tma)
    // (free-surface rows are redone after the sweep and Born steps a second field: those use the copy path)
    const bool fused = p2p && halo_fused_ok(a->halo, p) && !a->free_surface && !born;
    IsoFuse fz;
    if (fused && (rc = halo_fuse_desc(a->halo, p, fz))) return cleanup(rc);
    if (a->halo && nd == 3 && (a->x_m != 0 || a->x_M != p.a[0] - 2 * so - 1)) {
        set_error("b2_iso_forward: the decomposed dimension must be iterated over the whole slab "
                  "(x_m=%d, x_M=%d, %d owned planes)", a->x_m, a->x_M, p.a[0] - 2 * so);
        return cleanup(B2_ERR_INVALID);
    }
    if (a->halo) a->halo->reset_primed();           // first step of a call exchanges through NCCL
    const int dir = a->adjoint ? -1 : 1;
    // profile: [0] staging issued before the loop, [1] the loop (device clock)
    cudaEventRecord(ev_prof[0], stream());
    if (streamed) {
        if ((rc = iso_forward_streamed(a, p, g, u, staged_damp ? &damp : nullptr, staged_param ? &param : nullptr, src, rec,
                                       scalar_scale, dt2, damp_io, param_io)))
            return cleanup(rc);
        u_is_home = true;
        g_last_profile[4] = 1.0;
    }
    if (!streamed)
    for (int time = a->adjoint ? a->time_M : a->time_m; a->adjoint ? time >= a->time_m : time <= a->time_M;
         time += dir) {
        const int t0 = ((time % T) + T) % T;
        const int t1 = (((time + dir) % T) + T) % T;     // written
        const int t2 = (((time - dir) % T) + T) % T;     // the other time level read
        const bool per_step_events = timing && (step_index++ % sample_stride == 0) && se.used + 4 <= 16384;
        if (per_step_events) { se.next(); ++nsampled; }
        if (fused) {
            if ((rc = halo_step_iso_fused(a->halo, p, t0, t2, t1))) return cleanup(rc);
            // the injection below must reach the copies of my boundary planes in the neighbours' halos
            g.peer_lo = fz.peer_lo; g.peer_hi = fz.peer_hi;
            g.off_lo = (long long)t1 * fz.slot_lo + (long long)(p.o[0] + fz.n_lo) * p.sx;
            g.off_hi = (long long)t1 * fz.slot_hi + (long long)(p.o[0] - p.n[0]) * p.sx;
            g.nown = p.n[0]; g.pw = p.radius[0];
        } else if (a->halo) {
            if ((rc = halo_exchange_and_step_iso(a->halo, p, t0, t2, t1))) return cleanup(rc);
        } else {
            if ((rc = iso_step(p, t0, t2, t1, 0, p.n[0]))) return cleanup(rc);
        }
        if (a->free_surface && (rc = iso_fs_fix(p, t0, t2, t1, 0, p.n[0]))) return cleanup(rc);
        if (per_step_events) se.next();
        float *f1 = p.u + (size_t)t1 * p.slot_elems;
        if ((rc = launch_inject(src, g, f1, nullptr, time, p.param_kind, p.param, scalar_scale, dt2)))
            return cleanup(rc);
        if (fused) {
            // the sweep stored the boundary planes, the injection patched them: release the flags
            if ((rc = halo_fused_signal(a->halo, p.u))) return cleanup(rc);
            if (a->rec_toff && rec.present && (rc = halo_p2p_wait(a->halo))) return cleanup(rc);
        } else if (p2p) {
            // boundary planes of u[t1] are final: store them into the neighbours' halos, signal
            if ((rc = halo_p2p_publish(a->halo, p.u, nullptr, p.slot_elems, t1, (size_t)p.sx, p.o[0], p.n[0],
                                       halo_width_iso(p))))
                return cleanup(rc);
            // receivers that sample the just-written time level may touch halo cells: those
            // arrive with the neighbours' stores of this same step
            if (a->rec_toff && rec.present && (rc = halo_p2p_wait(a->halo))) return cleanup(rc);
        }
        if (born) {
            // eqn2 of the reference's Born operator comes after the source injection into u[t+1]
            if (a->halo) {
                if ((rc = halo_exchange_and_step_iso(a->halo, pU, t0, t2, t1))) return cleanup(rc);
            } else if ((rc = iso_step(pU, t0, t2, t1, 0, pU.n[0]))) {
                return cleanup(rc);
            }
            if ((rc = iso_born_source(p, t0, t2, t1, pU.u + (size_t)t1 * p.slot_elems, (const float *)borndm.d,
                                      (long long)borndm.size[1] * borndm.size[2], borndm.size[2],
                                      a->x_m + dmh, a->y_m + dmh, a->z_m + dmh)))
                return cleanup(rc);
            // the linearised field's boundary planes are final: publish them like u's
            if (a->halo && halo_p2p_active(a->halo, pU.u) &&
                (rc = halo_p2p_publish(a->halo, pU.u, nullptr, pU.slot_elems, t1, (size_t)pU.sx, pU.o[0], pU.n[0],
                                       halo_width_iso(pU))))
                return cleanup(rc);
        }
        if (a->snap && time % a->snap_factor == 0) {
            // internal dims: a 2-D grid is (1, x, y), its snapshots (nsnaps, x, y)
            const long long ssy = snap.size[nd];
            const long long ssx = nd == 3 ? (long long)snap.size[2] * snap.size[3] : 0;
            const size_t one = (size_t)(nd == 3 ? snap.size[1] : 1) * snap.size[nd - 1] * snap.size[nd];
            float *dst = nullptr;
            if (streamer.active) {
                if ((rc = streamer.acquire(&dst))) return cleanup(rc);
            } else {
                dst = (float *)snap.d + (size_t)(time / a->snap_factor) * one;
            }
            const int d0 = nd == 3 ? a->x_m + snap_h : 0;
            const int d1 = (nd == 3 ? a->y_m : a->x_m) + snap_h;
            const int d2 = (nd == 3 ? a->z_m : a->y_m) + snap_h;
            if ((rc = iso_snapshot(p, a->snap_toff ? t1 : t0, dst, ssx, ssy, d0, d1, d2))) return cleanup(rc);
            if (streamer.active) {
                const int dd[3] = {d0, d1, d2};
                if ((rc = streamer.release(time / a->snap_factor, &snap.size[1], dd, p.n))) return cleanup(rc);
            }
        }
        if (per_step_events) se.next();
        if (a->halo && !p2p && a->rec_toff && rec.present) {
            // NCCL path: receivers that sample the level just written may reach into the halo, which is only
            // exchanged at the start of the next step -- exchange it now (the peer paths wait on the flags instead)
            if ((rc = halo_exchange_slot(a->halo, born ? pU : p, t1))) return cleanup(rc);
        }
        const float *fr = (born ? pU.u : p.u) + (size_t)(a->rec_toff ? t1 : t0) * p.slot_elems;
        if ((rc = launch_interp(rec, g, fr, nullptr, time))) return cleanup(rc);
        if (staged_grad) {
            if (time < 0 || time >= usave.size[0]) {
                set_error("b2_iso_forward: time=%d outside the saved wavefield (nt=%d)", time, usave.size[0]);
                return cleanup(B2_ERR_INVALID);
            }
            ImgK ik;
            ik.grad = (float *)grad.d;
            ik.us = (const float *)usave.d + (size_t)time * p.slot_elems;
            ik.v0 = p.u + (size_t)t0 * p.slot_elems;
            ik.v1 = p.u + (size_t)t1 * p.slot_elems;
            ik.v2 = p.u + (size_t)t2 * p.slot_elems;
            ik.sx = p.sx; ik.sy = p.sy;
            ik.gsy = grad.size[2];
            ik.gsx = (long long)grad.size[1] * grad.size[2];
            const int gh = a->grad->hsize ? a->grad->hsize[0] : (grad.size[0] - (u.size[1] - 2 * so)) / 2;
            ik.n0 = p.n[0]; ik.n1 = p.n[1]; ik.n2 = p.n[2];
            ik.o0 = p.o[0]; ik.o1 = p.o[1]; ik.o2 = p.o[2];
            ik.g0 = a->x_m + gh; ik.g1 = a->y_m + gh; ik.g2 = a->z_m + gh;
            ik.inv_dt2 = 1.0f / (a->dt * a->dt);
            dim3 blk(64, 4, 1), grd((ik.n2 + 63) / 64, (ik.n1 + 3) / 4, (unsigned)(ik.n0 < 65535 ? ik.n0 : 65535));
            k_imaging<<<grd, blk, 0, stream()>>>(ik);
            count_launch();
            B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
        }
        if (per_step_events) se.next();
        if (a->errctl && ((time - a->time_m) % 100 == 99 || time == (a->adjoint ? a->time_m : a->time_M))) {
            bool bad = false;
            if ((rc = check_finite(f1, p.slot_elems, bad))) return cleanup(rc);
            if (bad) { set_error("NaN/Inf detected in u at time=%d", time); return cleanup(B2_ERR_NAN); }
        }
    }
    if (a->halo && (rc = halo_p2p_drain(a->halo))) return cleanup(rc);
    if (timing) cudaEventRecord(ev_tot[1], stream());
    cudaEventRecord(ev_prof[1], stream());
    if (streamed && a->errctl) {
        bool bad = false;
        const int tl = (((a->time_M + 1) % T) + T) % T;
        if ((rc = check_finite(p.u + (size_t)tl * p.slot_elems, p.slot_elems, bad))) return cleanup(rc);
        if (bad) { set_error("NaN/Inf detected in u at time=%d", a->time_M); return cleanup(B2_ERR_NAN); }
    }

    cudaError_t e = cudaStreamSynchronize(stream());
    if (e != cudaSuccess) {
        set_error("b2_iso_forward: device error: %s", cudaGetErrorString(e));
        return cleanup(B2_ERR_LAUNCH);
    }
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev_prof[0], ev_prof[1]) == cudaSuccess) g_last_profile[1] = ms;
        g_last_profile[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - call_t0).count()
                            - g_last_profile[1];          // everything before the loop (staging in), host clock
    }
    if (timing) {
        float tot = 0.f;
        cudaEventElapsedTime(&tot, ev_tot[0], ev_tot[1]);
        double s1 = 0, s2 = 0;
        for (size_t i = 0; i + 3 < se.used; i += 4) {
            float ms;
            if (cudaEventElapsedTime(&ms, se.ev[i + 1], se.ev[i + 2]) == cudaSuccess) s1 += ms;
            if (cudaEventElapsedTime(&ms, se.ev[i + 2], se.ev[i + 3]) == cudaSuccess) s2 += ms;
        }
        if (nsampled > 0) {
            const double up = (double)nsteps / nsampled;
            s1 *= up;
            s2 *= up;
        }
        if (s1 + s2 > tot) { const double f = tot / (s1 + s2 + 1e-30); s1 *= f * 0.5; s2 *= f * 0.5; }
        a->timers->section0 += (tot - s1 - s2) * 1e-3;
        a->timers->section1 += s1 * 1e-3;
        a->timers->section2 += s2 * 1e-3;
    }
    return cleanup(B2_OK);
}
