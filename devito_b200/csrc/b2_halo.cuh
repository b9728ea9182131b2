// Halo exchange under x-slab decomposition (internal interface).
#pragma once
#include "b2_common.cuh"
#include "b2_iso.cuh"
#include "b2_tti.cuh"
#include <vector>

struct b2_halo_ctx {
    int rank = 0, nranks = 1, deviceid = 0;
    void *nccl_dl = nullptr;
    void *comm = nullptr;                 // ncclComm_t
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_comm = nullptr;
    double seconds = 0.0;                 // accumulated exchange time (when timed)
    // peer-memory path
    bool p2p = false;
    int *flags_local = nullptr;           // [0] written by the left neighbour, [1] by the right one
    int *flag_left_remote = nullptr;      // slot in the LEFT neighbour's buffer that we signal
    int *flag_right_remote = nullptr;
    int step = 0;                         // monotonic count of completed p2p steps
    bool p2p_primed = false;              // halos of the current u[t0] were stored by the peers
    // opt-in (B2_P2P_ASYNC=1, not yet validated on hardware): peer stores + signal go to a
    // high-priority side stream so that they overlap the next step's interior update
    bool p2p_async = false;
    cudaStream_t push_stream = nullptr;
    cudaEvent_t ev_final = nullptr, ev_push = nullptr;
    bool push_pending = false;
    // `primed`: the halos of this field's current time level were stored by the peers (false at the start of
    // a call: the first step exchanges through NCCL). Per field: Born modelling steps two wavefields.
    struct Reg { void *local, *left, *right; int n_left, n_right; bool primed; };
    std::vector<Reg> regs;
    Reg *find(const void *local) {
        for (Reg &r : regs) if (r.local == local) return &r;
        return nullptr;
    }
    void reset_primed() { for (Reg &r : regs) r.primed = false; p2p_primed = false; }
    bool primed(const void *local) { Reg *r = find(local); return r && r->primed; }
    void set_primed(const void *local) { if (Reg *r = find(local)) r->primed = true; p2p_primed = true; }
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

// peer-memory path helpers (see b2_halo.cu)
int halo_p2p_active(b2_halo_ctx *ctx, const void *base);
int halo_p2p_wait(b2_halo_ctx *ctx);     // block the stream until both neighbours signalled the current step
int halo_p2p_publish(b2_halo_ctx *ctx, const float *f0, const float *f1, size_t slot_elems, int slot1,
                     size_t plane, int lo, int n, int width);
// make the main stream wait for an outstanding asynchronous push (no-op otherwise); called before the
// boundary strips of the next step and once after the time loop
int halo_p2p_drain(b2_halo_ctx *ctx);

// Fused path (isotropic TMA sweep): the kernel stores its boundary planes into the neighbours' halos and
// acquires their flags itself; halo_fused_signal releases this rank's flags after the step's injection.
bool halo_fused_ok(b2_halo_ctx *ctx, const IsoPlan &p);
int halo_fuse_desc(b2_halo_ctx *ctx, const IsoPlan &p, IsoFuse &f);
int halo_step_iso_fused(b2_halo_ctx *ctx, const IsoPlan &p, int t0, int t2, int t1);
int halo_fused_signal(b2_halo_ctx *ctx, const void *field);
int halo_width_iso(const IsoPlan &p);
int halo_exchange_initial(b2_halo_ctx *ctx, const IsoPlan &p, int t0);
int halo_exchange_slot(b2_halo_ctx *ctx, const IsoPlan &p, int slot);

// Same for the coupled TTI fields: u and v boundary planes travel in one NCCL group.
int halo_exchange_and_step_tti(b2_halo_ctx *ctx, const TtiPlan &p, int t0, int t2, int t1);

}  // namespace b2
