// TTI centred forward kernels for sm_100a, scalar Thomsen parameters.
//
// Numerical spec (reference: examples/seismic/tti/operators.py:65-104 `Gzz_centered`,
// :146-183 `Gh_centered`, :186-247 `kernel_centered`, :12-39 `second_order_stencil`; the
// operation order below follows the code the reference generates for `ForwardTTI`):
//
//   D+_d f (p) = sum_j w1_d[j] f(p + (j - R/2 + 1) e_d)     half-node derivative at p + h/2
//   D-_d g (p) = sum_j w1_d[j] g(p + (j - R/2)     e_d)     half-node derivative at p - h/2
//   Gz(f)  = sin(th)cos(ph) D+_x f + sin(th)sin(ph) D+_y f + cos(th) D+_z f
//   Gzz(f) = D-_z(cos(th) Gz f) + D-_x(sin(th)cos(ph) Gz f) + D-_y(sin(th)sin(ph) Gz f)
//   H0 = (1+2 eps) (lap(u) - Gzz(u)) + sqrt(1+2 delta) Gzz(v)
//   Hz = sqrt(1+2 delta) (lap(u) - Gzz(u)) + Gzz(v)
//   u+ = ( m/dt^2 (2u - u-) + damp/dt u + H0 ) / ( m/dt^2 + damp/dt ),  v+ likewise with Hz
//
// Kernel 1 (two-pass, any even radius): pass A writes Gz(u), Gz(v) to scratch, pass B
// applies the outer derivatives + Laplacian + time update.
#include "b2_tti.cuh"
#include <algorithm>

namespace b2 {

struct TtiK {
    const float *__restrict__ u0;
    const float *__restrict__ v0;
    const float *__restrict__ um;
    const float *__restrict__ vm;
    float *__restrict__ u1;
    float *__restrict__ v1;
    const float *__restrict__ damp;
    float *__restrict__ gzu;
    float *__restrict__ gzv;
    long long sx, sy;
    int n0, n1, n2;
    int o0, o1, o2;
    int R;
    float m_dt2, inv_dt;
    float cx, cy, cz;          // sin(th)cos(ph), sin(th)sin(ph), cos(th)
    float e2, sd;              // 1+2eps, sqrt(1+2delta)
    float w2[3][B2_MAX_RADIUS + 1];
    float w1[3][B2_MAX_RADIUS];
};

// pass A: Gz over the box extended by [-R/2, R/2-1]
__global__ void __launch_bounds__(256) k_tti_gz(TtiK k) {
    const int h = k.R / 2;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 + k.R || y >= k.n1 + k.R) return;
    for (int x = blockIdx.z; x < k.n0 + k.R; x += gridDim.z) {
        const long long idx = (long long)(k.o0 + x - h) * k.sx + (long long)(k.o1 + y - h) * k.sy + (k.o2 + z - h);
        float dxu = 0.f, dyu = 0.f, dzu = 0.f, dxv = 0.f, dyv = 0.f, dzv = 0.f;
        for (int j = 0; j < k.R; ++j) {
            const int off = j - h + 1;
            dxu = fmaf(k.w1[0][j], k.u0[idx + off * k.sx], dxu);
            dyu = fmaf(k.w1[1][j], k.u0[idx + off * k.sy], dyu);
            dzu = fmaf(k.w1[2][j], k.u0[idx + off], dzu);
            dxv = fmaf(k.w1[0][j], k.v0[idx + off * k.sx], dxv);
            dyv = fmaf(k.w1[1][j], k.v0[idx + off * k.sy], dyv);
            dzv = fmaf(k.w1[2][j], k.v0[idx + off], dzv);
        }
        k.gzu[idx] = k.cx * dxu + k.cy * dyu + k.cz * dzu;
        k.gzv[idx] = k.cx * dxv + k.cy * dyv + k.cz * dzv;
    }
}

__global__ void __launch_bounds__(256) k_tti_update(TtiK k) {
    const int h = k.R / 2;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) {
        const long long idx = (long long)(k.o0 + x) * k.sx + (long long)(k.o1 + y) * k.sy + (k.o2 + z);
        const float uc = k.u0[idx], vc = k.v0[idx];
        float lap = (k.w2[0][0] + k.w2[1][0] + k.w2[2][0]) * uc;
        for (int i = 1; i <= k.R; ++i) {
            lap = fmaf(k.w2[0][i], k.u0[idx - i * k.sx] + k.u0[idx + i * k.sx], lap);
            lap = fmaf(k.w2[1][i], k.u0[idx - i * k.sy] + k.u0[idx + i * k.sy], lap);
            lap = fmaf(k.w2[2][i], k.u0[idx - i] + k.u0[idx + i], lap);
        }
        float gu = 0.f, gv = 0.f;
        for (int j = 0; j < k.R; ++j) {
            const int off = j - h;
            gu = fmaf(k.cx * k.w1[0][j], k.gzu[idx + off * k.sx], gu);
            gu = fmaf(k.cy * k.w1[1][j], k.gzu[idx + off * k.sy], gu);
            gu = fmaf(k.cz * k.w1[2][j], k.gzu[idx + off], gu);
            gv = fmaf(k.cx * k.w1[0][j], k.gzv[idx + off * k.sx], gv);
            gv = fmaf(k.cy * k.w1[1][j], k.gzv[idx + off * k.sy], gv);
            gv = fmaf(k.cz * k.w1[2][j], k.gzv[idx + off], gv);
        }
        const float gh = lap - gu;                  // Gxx + Gyy applied to u
        const float H0 = k.e2 * gh + k.sd * gv;
        const float Hz = k.sd * gh + gv;
        const float d = k.damp ? k.damp[idx] * k.inv_dt : 0.f;
        const float den = k.m_dt2 + d;
        k.u1[idx] = (k.m_dt2 * (2.f * uc - k.um[idx]) + d * uc + H0) / den;
        k.v1[idx] = (k.m_dt2 * (2.f * vc - k.vm[idx]) + d * vc + Hz) / den;
    }
}

int tti_plan_init(TtiPlan &p, int kernel) {
    p.kernel = kernel;
    if (p.R % 2 != 0 || p.R < 2 || p.R > B2_MAX_RADIUS) {
        set_error("tti: radius %d unsupported (space_order must be a multiple of 4, <= 16)", p.R);
        return B2_ERR_INVALID;
    }
    const size_t bytes = p.slot_elems * sizeof(float);
    B2_CUDA(cudaMalloc(&p.gzu, bytes), B2_ERR_MEMORY);
    B2_CUDA(cudaMalloc(&p.gzv, bytes), B2_ERR_MEMORY);
    return B2_OK;
}

void tti_plan_free(TtiPlan &p) {
    if (p.gzu) cudaFree(p.gzu);
    if (p.gzv) cudaFree(p.gzv);
    p.gzu = p.gzv = nullptr;
}

int tti_step(const TtiPlan &p, int slot0, int slotm, int slot1, int xlo, int xcount) {
    if (xcount <= 0) return B2_OK;
    TtiK k;
    k.u0 = p.u + (size_t)slot0 * p.slot_elems;
    k.v0 = p.v + (size_t)slot0 * p.slot_elems;
    k.um = p.u + (size_t)slotm * p.slot_elems;
    k.vm = p.v + (size_t)slotm * p.slot_elems;
    k.u1 = p.u + (size_t)slot1 * p.slot_elems;
    k.v1 = p.v + (size_t)slot1 * p.slot_elems;
    k.damp = p.damp;
    k.gzu = p.gzu;
    k.gzv = p.gzv;
    k.sx = p.sx;
    k.sy = p.sy;
    k.n0 = xcount;
    k.n1 = p.n[1];
    k.n2 = p.n[2];
    k.o0 = p.o[0] + xlo;
    k.o1 = p.o[1];
    k.o2 = p.o[2];
    k.R = p.R;
    k.inv_dt = 1.0f / p.dt;
    k.m_dt2 = (1.0f / (p.vp * p.vp)) * (1.0f / (p.dt * p.dt));
    const float st = sinf(p.theta), ct = cosf(p.theta), sp = sinf(p.phi), cp = cosf(p.phi);
    k.cx = st * cp;
    k.cy = st * sp;
    k.cz = ct;
    k.e2 = 1.0f + 2.0f * p.epsilon;
    k.sd = sqrtf(1.0f + 2.0f * p.delta);
    memcpy(k.w2, p.w2, sizeof(k.w2));
    memcpy(k.w1, p.w1, sizeof(k.w1));
    dim3 block(64, 4, 1);
    dim3 gridA((k.n2 + k.R + 63) / 64, (k.n1 + k.R + 3) / 4, (unsigned)std::min(xcount + k.R, 65535));
    timing_begin();
    k_tti_gz<<<gridA, block, 0, stream()>>>(k);
    count_launch();
    dim3 gridB((k.n2 + 63) / 64, (k.n1 + 3) / 4, (unsigned)std::min(xcount, 65535));
    k_tti_update<<<gridB, block, 0, stream()>>>(k);
    timing_end();
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

}  // namespace b2
