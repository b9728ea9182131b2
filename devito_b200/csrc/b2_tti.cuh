// TTI (tilted transverse isotropy) centred-kernel stencils — internal interface.
#pragma once
#include "b2_common.cuh"

namespace b2 {

struct TtiPlan {
    int R = 4;                   // radius = space_order/2 (even)
    int so = 8;
    int a[3] = {1, 1, 1};
    int n[3] = {1, 1, 1};
    int o[3] = {0, 0, 0};
    long long sx = 0, sy = 0;
    size_t slot_elems = 0;
    int tsize = 3;
    float *u = nullptr, *v = nullptr;
    const float *damp = nullptr;
    float vp = 1.f, dt = 1.f, epsilon = 0.f, delta = 0.f, theta = 0.f, phi = 0.f;
    float w2[3][B2_MAX_RADIUS + 1] = {};
    float w1[3][B2_MAX_RADIUS] = {};
    // scratch (two-pass kernel): Gz(u), Gz(v)
    float *gzu = nullptr, *gzv = nullptr;
    int kernel = 0;
    // fused single-pass kernel
    bool use_fused = false;
    bool arr_fused = false;      // fused kernel with per-point parameter tables (k_tti_fused<.., ARR = true>)
    CUtensorMap tm_u, tm_v;
    CUtensorMap tm_cx, tm_cy, tm_cz;   // factor tables (arr_ct)
    int arr_ct = 0;              // ... 1: stage-A factor tiles staged through shared memory by TMA; 2: stage B reads them too
    float *coefA = nullptr;
    // array-valued parameters (device pointers, nullptr -> scalar)
    const float *vp_a = nullptr, *eps_a = nullptr, *delta_a = nullptr, *theta_a = nullptr, *phi_a = nullptr;
    // tabulated per-point coefficients (library scratch) when any parameter is an array
    float *tCx = nullptr, *tCy = nullptr, *tCz = nullptr, *tE2 = nullptr, *tSD = nullptr, *tMD = nullptr;
    bool has_arrays = false;
};

int tti_plan_init(TtiPlan &p, int kernel);
void tti_plan_free(TtiPlan &p);
// one time step over x in [xlo, xlo+xcount)
int tti_step(const TtiPlan &p, int slot0, int slotm, int slot1, int xlo, int xcount);

}  // namespace b2
