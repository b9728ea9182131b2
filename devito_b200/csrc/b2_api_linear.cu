// C-ABI entry point b2_linear_forward: explicit time stepping of ONE field with a constant-coefficient
// linear stencil — the generated `Kernel` of operators like the reference's 2-D diffusion example
// (examples/cfd/example_diffusion.py:120-133; BASELINE config 1):
//
//   for time in [time_m, time_M]:  f[(time + wshift) % T][p] = sum_k coef_k * f[(time + tshift_k) % T][p + off_k]
//
// One thread per point (k_linear), taps in the kernel parameter block. This is the plumbing-size path
// (512^2 in BASELINE config 1), not a tuned sweep: the hot wave-propagation schemes have their own kernels.
#include "b2_common.cuh"
#include "b2_iso_point.cuh"
#include <algorithm>

using namespace b2;

namespace b2 {

__global__ void __launch_bounds__(256) k_linear(LinK k) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) linear_point(k, x, y, z);
}

}  // namespace b2

extern "C" int b2_linear_forward(const struct b2_linear_args *a) {
    if (!a || !a->f || !a->taps) { set_error("b2_linear_forward: NULL args"); return B2_ERR_INVALID; }
    if (a->ndim < 1 || a->ndim > 3) { set_error("b2_linear_forward: ndim must be 1..3"); return B2_ERR_INVALID; }
    if (a->ntaps < 1 || a->ntaps > B2_MAX_TAPS) {
        set_error("b2_linear_forward: %d taps (1..%d supported)", a->ntaps, B2_MAX_TAPS);
        return B2_ERR_INVALID;
    }
    if (a->wshift != 1 && a->wshift != -1) { set_error("b2_linear_forward: wshift must be +1 or -1"); return B2_ERR_INVALID; }
    if (a->time_M < a->time_m) return B2_OK;
    std::lock_guard<std::mutex> api_lock(api_mutex());
    if (int rc0 = use_device(a->deviceid)) return rc0;

    const int nd = a->ndim, h = a->halo;
    DevArray f;
    int rc = stage_in(a->f, nd + 1, f, true);
    if (rc) return rc;
    auto cleanup = [&](int code) {
        const int r = stage_out(f, code == B2_OK);
        return code != B2_OK ? code : r;
    };
    // internal 3-dim convention: leading dummy dims of extent 1 for 1-D / 2-D fields
    int alloc[3] = {1, 1, 1}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, hal[3] = {0, 0, 0};
    const int lo_in[3] = {a->x_m, a->y_m, a->z_m}, hi_in[3] = {a->x_M, a->y_M, a->z_M};
    for (int d = 0; d < nd; ++d) {
        const int di = 3 - nd + d;
        alloc[di] = f.size[d + 1];
        lo[di] = lo_in[d];
        hi[di] = hi_in[d];
        hal[di] = h;
    }
    const int T = f.size[0];
    LinK k;
    k.sy = alloc[2];
    k.sx = (long long)alloc[1] * alloc[2];
    const size_t slot = (size_t)alloc[0] * alloc[1] * alloc[2];
    k.n0 = hi[0] - lo[0] + 1; k.n1 = hi[1] - lo[1] + 1; k.n2 = hi[2] - lo[2] + 1;
    k.o0 = lo[0] + hal[0]; k.o1 = lo[1] + hal[1]; k.o2 = lo[2] + hal[2];
    if (k.n0 <= 0 || k.n1 <= 0 || k.n2 <= 0) return cleanup(B2_OK);
    k.ntaps = a->ntaps;
    int shifts[4], nshift = 0;
    for (int i = 0; i < a->ntaps; ++i) {
        const b2_tap &t = a->taps[i];
        if (t.tshift == a->wshift) { set_error("b2_linear_forward: a tap reads the level being written"); return cleanup(B2_ERR_INVALID); }
        int s = -1;
        for (int j = 0; j < nshift; ++j) if (shifts[j] == t.tshift) s = j;
        if (s < 0) {
            if (nshift == 4) { set_error("b2_linear_forward: more than 4 time levels read"); return cleanup(B2_ERR_INVALID); }
            shifts[nshift] = t.tshift;
            s = nshift++;
        }
        k.sel[i] = s;
        long long delta = 0;
        for (int d = 0; d < nd; ++d) {
            const int di = 3 - nd + d, o = t.off[d];
            // the read must stay inside the allocated array
            if (lo[di] + hal[di] + o < 0 || hi[di] + hal[di] + o >= alloc[di]) {
                set_error("b2_linear_forward: tap offset %d on dim %d leaves the allocated array", o, d);
                return cleanup(B2_ERR_INVALID);
            }
            delta += (long long)o * (di == 0 ? k.sx : di == 1 ? k.sy : 1);
        }
        k.delta[i] = delta;
        k.coef[i] = t.coef;
    }
    if (T <= nshift) {   // every level read and the one written must be distinct slots
        set_error("b2_linear_forward: %d time slots cannot hold %d levels", T, nshift + 1);
        return cleanup(B2_ERR_INVALID);
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (a->timers) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream());
    }
    dim3 block(64, 4, 1);
    dim3 grid((k.n2 + 63) / 64, (k.n1 + 3) / 4, (unsigned)std::min(k.n0, 65535));
    float *base = (float *)f.d;
    auto slot_of = [&](int time, int shift) { return (((time + shift) % T) + T) % T; };
    // an update of `f.backward` (wshift = -1) runs from time_M down to time_m, like the reference's
    // backward-in-time loop (devito/ir/support/space.py: Backward direction)
    const int dirn = a->wshift;
    for (int time = dirn > 0 ? a->time_m : a->time_M; dirn > 0 ? time <= a->time_M : time >= a->time_m; time += dirn) {
        k.out = base + (size_t)slot_of(time, a->wshift) * slot;
        for (int j = 0; j < nshift; ++j) k.lvl[j] = base + (size_t)slot_of(time, shifts[j]) * slot;
        for (int j = nshift; j < 4; ++j) k.lvl[j] = k.lvl[0];
        k_linear<<<grid, block, 0, stream()>>>(k);
        count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (a->timers) cudaEventRecord(e1, stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream());
    if (a->timers) {
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) a->timers->section0 += ms * 1e-3;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    if (e != cudaSuccess) { set_error("b2_linear_forward: %s", cudaGetErrorString(e)); return cleanup(B2_ERR_LAUNCH); }
    return cleanup(B2_OK);
}
