// Halo exchange under x-slab decomposition (internal interface).
#pragma once
#include "b2_common.cuh"
#include "b2_iso.cuh"
#include "b2_tti.cuh"

struct b2_halo_ctx {
    int rank = 0, nranks = 1, deviceid = 0;
    void *nccl_dl = nullptr;
    void *comm = nullptr;                 // ncclComm_t
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_comm = nullptr;
    double seconds = 0.0;                 // accumulated exchange time (when timed)
};

namespace b2 {

// enqueue the exchange of `width` planes of `field` (one time slot, base pointer given) on
// the comm stream; plane_elems = elements of one yz-plane (incl. halos); lo = array x index
// of the first owned plane; n = owned planes.
int halo_enqueue(b2_halo_ctx *ctx, float *slot_base, size_t plane_elems, int lo, int n, int width);

// One isotropic time step with the exchange of u[t0] overlapped with the interior update:
//   comm stream : send/recv boundary planes of u[t0]
//   main stream : interior x in [R, n-R)  ->  wait comm  ->  strips [0,R) and [n-R,n)
int halo_exchange_and_step_iso(b2_halo_ctx *ctx, const IsoPlan &p, int t0, int t2, int t1);

// Same for the coupled TTI fields: u and v boundary planes travel in one NCCL group.
int halo_exchange_and_step_tti(b2_halo_ctx *ctx, const TtiPlan &p, int t0, int t2, int t1);

}  // namespace b2
