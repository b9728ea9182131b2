// Isotropic acoustic time-update kernels for sm_100a.
//
// Numerical spec (reference: examples/seismic/acoustic/operators.py:71-107 `iso_stencil`,
// `laplacian` :50-68; generated form printed by the reference's own code generator):
//
//   u[t+1] = ( m/dt^2 (2 u[t] - u[t-1]) + damp/dt u[t] + sum_d sum_k w_d[k] u[t][..+k..] )
//            / ( m/dt^2 + damp/dt )
//
// with w_d[k] = finite_diff_weights(2, ...)/h_d^2 (devito/finite_differences/tools.py:231-236).
//
// Two kernels:
//   k_iso_generic : one thread per point, any radius <= 8, 2-D or 3-D, any alignment.
//   k_iso_tma     : 2.5-D sweep. A CTA owns a (TY x TZ) yz-tile and marches along x.
//                   A producer warp streams haloed yz-planes of u[t] (and the matching tiles
//                   of u[t-1], damp, vp|m) into shared-memory rings with TMA
//                   (cp.async.bulk.tensor) signalled on mbarriers; consumer threads own four
//                   consecutive z points (float4), keep the x-direction neighbours in a
//                   register queue, read y/z neighbours from the shared plane, and write
//                   u[t+1] with 16-byte coalesced stores. No tensor cores: the update is
//                   HBM-bandwidth bound (16 B/point).
#include "b2_iso.cuh"
#include "b2_iso_point.cuh"
#include "b2_ptx.cuh"
#include <cstdlib>
#include <algorithm>

namespace b2 {

// ------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------
// point forms (shared with the CPU emulation test): b2_iso_point.cuh
__global__ void __launch_bounds__(256) k_iso_generic(IsoGK k) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) iso_point(k, x, y, z);
}

// OT4 first pass over the box grown by the radius (n*/o* already describe the grown box)
__global__ void __launch_bounds__(256) k_ot4_w(IsoGK k) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) ot4_w_point(k, x, y, z);
}

// Time-subsampled snapshot of the iteration box (b2_iso_point.cuh::snapshot_point)
__global__ void __launch_bounds__(256) k_snapshot(SnapK k) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) snapshot_point(k, x, y, z);
}

int iso_snapshot(const IsoPlan &p, int slot, float *snap, long long dsx, long long dsy, int d0, int d1, int d2) {
    SnapK k;
    k.src = p.u + (size_t)slot * p.slot_elems;
    k.dst = snap;
    k.sx = p.sx; k.sy = p.sy;
    k.dsx = dsx; k.dsy = dsy;
    k.n0 = p.n[0]; k.n1 = p.n[1]; k.n2 = p.n[2];
    k.o0 = p.o[0]; k.o1 = p.o[1]; k.o2 = p.o[2];
    k.d0 = d0; k.d1 = d1; k.d2 = d2;
    dim3 block(64, 4, 1);
    dim3 grid((k.n2 + 63) / 64, (k.n1 + 3) / 4, (unsigned)std::min(k.n0, 65535));
    k_snapshot<<<grid, block, 0, stream()>>>(k);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

// Born source term added to the freshly updated linearised field (b2_iso_point.cuh::born_src_point)
__global__ void __launch_bounds__(256) k_born_src(IsoGK k) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (z >= k.n2 || y >= k.n1) return;
    for (int x = blockIdx.z; x < k.n0; x += gridDim.z) born_src_point(k, x, y, z);
}

// Free surface: the sweep kernels compute every row with the plain stencil; this kernel then REDOES
// the rows z <= radius — the only ones with a vertical tap that reaches the surface (z - k == 0, which
// contributes 0 in the reference, not u[0]: u[0] is non-zero right after a source deposited into the
// surface row) or goes above it — and clears the surface row. It touches (radius+1)/n2 of the grid
// (0.5 % at 1024^3, so = 8). The halo above the surface is left alone: sinc sources/receivers deposit
// into / sample it in the reference too.
__global__ void __launch_bounds__(256) k_iso_fs_fix(IsoGK k) {
    const int z = threadIdx.x;                                   // 0 .. r2 (blockDim.x == 16 > B2_MAX_RADIUS)
    const int y = blockIdx.x * blockDim.y + threadIdx.y;
    if (z > k.r2 || y >= k.n1) return;
    for (int x = blockIdx.y; x < k.n0; x += gridDim.y) iso_fs_point(k, x, y, z);
}

// ------------------------------------------------------------------------------------------
// TMA 2.5-D kernel
// ------------------------------------------------------------------------------------------
// The update is evaluated in the algebraically identical division-free form
//     u+ = u + A (m/dt^2 (u - u-) + lap)            scalar m      (A = 1/(m/dt^2 + damp/dt))
//     u+ = u + B (u - u-) + A lap                   array vp|m    (B = m/dt^2 * A)
// with A (and B) tabulated once per call by k_iso_coef: the IEEE division, which cost ~1/5 of
// the issued instructions of the first version of this kernel (profiles/r1a_*), leaves the
// inner loop, and A replaces damp (B replaces vp|m) in the streamed arrays, so the HBM traffic
// per point is unchanged (16 B, or 20 B with an array parameter).
template <int R>
struct IsoTK {
    float *__restrict__ u1;      // base of the output time slot
    long long sx, sy;
    int ny, nz;                  // iteration extents (for masking)
    int ox, oy, oz;
    int xlo, xcount, lx;
    int ntz, nty;
    int slot0, slotm;
    float m_dt2;
    float wx[R + 1], wy[R + 1], wz[R + 1];
    // k_iso_tma2 (two y rows per thread): u[t-1] and the coefficient tables are read straight from global memory
    // (coalesced 16-byte loads, prefetched into registers), packed fp32x2 weights {w, w}
    const float *um, *cA, *cB;
    float2 p_wx[R + 1], p_wy[R + 1], p_mdt2, p_wc;
    // ---- x-slab decomposition, halo step fused into the sweep (all zero / null on a single device) ----
    // The CTAs that produce a boundary plane also store it into the neighbour GPU's halo through the
    // CUDA-IPC mapped peer pointer (16-byte stores over NVLink), and the producer lane of a CTA that
    // needs a halo plane of u[t] first acquires the flag the neighbour released after its previous
    // step. Chunk 0 marches from its high end DOWN to x = 0 when a low neighbour exists, so every CTA
    // touches the neighbour's planes at the very END of its march: by then the flag has long been set.
    int back0;                    // chunk 0 marches backwards
    int nown;                     // planes owned by this rank: peer planes are [0, pw) and [nown - pw, nown)
    int pw;                       // planes published per side (= R)
    float *peer_lo, *peer_hi;     // neighbour field bases (NULL: physical boundary)
    long long off_lo, off_hi;     // element offset of "my plane 0" inside the neighbour's array (output slot)
    const int *flag_lo, *flag_hi; // local flags the neighbours release (NULL: nothing to wait for)
    int want;                     // flag value that says "halos of u[t] are in place" (< 0: already ensured)
};

constexpr int ceil4(int v) { return (v + 3) / 4 * 4; }
constexpr int align32f(int v) { return (v + 31) / 32 * 32; }   // 128-byte multiples in floats

// Ring layout: the u[t] planes live in a ring of exactly Q = 2R+1 stages, so that inside the
// loop body (unrolled Q times) every shared-memory address is a compile-time offset; plane j
// sits in stage j % Q, and the R planes beyond the one being read for output are the prefetch
// depth. The (u[t-1], A[, B]) tiles of output step s use a second, shorter ring of NS >= R+1
// stages indexed at run time (two loads per iteration only). One full/empty mbarrier pair per
// u-stage covers both: the tiles of step s = j-2R travel with plane j.
template <int R, int TY, int TZ4, int PK>
struct IsoTmaCfg {
    static constexpr int RZ = ceil4(R);
    static constexpr int TZ = 4 * TZ4;
    static constexpr int BY = TY + 2 * R;
    static constexpr int BZ = TZ + 2 * RZ;
    static constexpr int Q = 2 * R + 1;
    static constexpr int NU = Q;
    static constexpr int NS = R + 1;
    static constexpr int PLANE = align32f(BY * BZ);
    static constexpr int TILE = TY * TZ;
    static constexpr int NCW = TY * TZ4 / 32;          // consumer warps
    static constexpr int NTILES = 2 + (PK != B2_PARAM_SCALAR ? 1 : 0);   // prev, A, [B]
    static constexpr size_t SMEM =
        (size_t)(NU * PLANE + NS * TILE * NTILES) * 4 + 2 * NU * 8 + 128;
};

__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) {
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ void f4fma(float4 &acc, float w, float4 v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

template <int R, int TY, int TZ4, int PK>
__global__ void __launch_bounds__(TY *TZ4 + 32, 1)
k_iso_tma(const __grid_constant__ CUtensorMap tm_uh, const __grid_constant__ CUtensorMap tm_uc,
          const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
          const IsoTK<R> k) {
    using C = IsoTmaCfg<R, TY, TZ4, PK>;
    constexpr int RZ = C::RZ, TZ = C::TZ, BZ = C::BZ, NU = C::NU, NS = C::NS, Q = C::Q;
    constexpr int PLANE = C::PLANE, TILE = C::TILE, NCW = C::NCW;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_u = reinterpret_cast<float *>(smem_raw);
    float *s_prev = s_u + NU * PLANE;
    float *s_a = s_prev + NS * TILE;
    float *s_b = s_a + NS * TILE;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_b + (PK != B2_PARAM_SCALAR ? NS * TILE : 0));
    uint64_t *empty = full + NU;

    int b = blockIdx.x;
    const int iz = b % k.ntz;
    b /= k.ntz;
    const int iy = b % k.nty;
    const int ix = b / k.nty;
    const int z0 = iz * TZ, y0 = iy * TY;
    const int xs = k.xlo + ix * k.lx;
    const int xe = min(xs + k.lx, k.xlo + k.xcount);
    const int NP = (xe - xs) + 2 * R;
    // plane j of the march sits at x = xb + dx * j (relative to the iteration origin); the output of
    // iteration j is the plane x = xb + dx * (j - R)
    const bool back = (ix == 0) && k.back0;
    const int dx = back ? -1 : 1;
    const int xb = back ? xe - 1 + R : xs - R;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NU; ++i) {
            b2ptx::mbar_init(&full[i], 1);
            b2ptx::mbar_init(&empty[i], NCW);
        }
        b2ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------- producer warp: one lane drives the TMA engine ----------------
        if (lane == 0) {
            b2ptx::tma_prefetch_desc(&tm_uh);
            b2ptx::tma_prefetch_desc(&tm_uc);
            b2ptx::tma_prefetch_desc(&tm_a);
            if (PK != B2_PARAM_SCALAR) b2ptx::tma_prefetch_desc(&tm_b);
            int slot = 0, round = 0, ss = 0;
            bool need_lo = k.want >= 0 && k.flag_lo != nullptr, need_hi = k.want >= 0 && k.flag_hi != nullptr;
            for (int j = 0; j < NP; ++j) {
                if (round > 0) b2ptx::mbar_wait(&empty[slot], (round - 1) & 1);
                const int s = j - 2 * R;
                const int xj = xb + dx * j;
                // a plane the neighbour GPU owns: it stored it into our halo at the end of its previous
                // step and then released the flag; acquire it once, then let the TMA engine read
                if (need_lo && xj < 0) {
                    while (b2ptx::ld_acquire_sys(k.flag_lo) < k.want) __nanosleep(40);
                    b2ptx::fence_proxy_async_global();
                    need_lo = false;
                }
                if (need_hi && xj >= k.nown) {
                    while (b2ptx::ld_acquire_sys(k.flag_hi) < k.want) __nanosleep(40);
                    b2ptx::fence_proxy_async_global();
                    need_hi = false;
                }
                uint32_t bytes = C::BY * BZ * 4;
                if (s >= 0) bytes += TILE * 4 * C::NTILES;
                b2ptx::mbar_arrive_expect_tx(&full[slot], bytes);
                b2ptx::tma_load_4d(s_u + slot * PLANE, &tm_uh, &full[slot], k.oz + z0 - RZ,
                                   k.oy + y0 - R, k.ox + xj, k.slot0);
                if (s >= 0) {
                    const int xo = xj - dx * R;
                    b2ptx::tma_load_4d(s_prev + ss * TILE, &tm_uc, &full[slot], k.oz + z0,
                                       k.oy + y0, k.ox + xo, k.slotm);
                    b2ptx::tma_load_3d(s_a + ss * TILE, &tm_a, &full[slot], k.oz + z0, k.oy + y0,
                                       k.ox + xo);
                    if (PK != B2_PARAM_SCALAR)
                        b2ptx::tma_load_3d(s_b + ss * TILE, &tm_b, &full[slot], k.oz + z0,
                                           k.oy + y0, k.ox + xo);
                    if (++ss == NS) ss = 0;
                }
                if (++slot == NU) { slot = 0; ++round; }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int ty = tid / TZ4, tz4 = tid % TZ4;
    const int gy = y0 + ty, gz = z0 + 4 * tz4;
    const bool yok = gy < k.ny;
    const int zcnt = yok ? min(max(k.nz - gz, 0), 4) : 0;
    const float *my_col = s_u + (ty + R) * BZ + RZ + 4 * tz4;
    const float *my_prev = s_prev + ty * TZ + 4 * tz4;
    const long long rowoff = (long long)(k.oy + gy) * k.sy + (k.oz + gz);
    float *outp = k.u1 + (long long)(k.ox + xb - dx * R) * k.sx + rowoff;
    const long long osx = dx * k.sx;
    const bool fuse = (k.peer_lo != nullptr) | (k.peer_hi != nullptr);

    float4 q[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) q[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const float wc = k.wx[0] + k.wy[0] + k.wz[0];
    uint32_t par = 0;               // parity of the u-ring round
    int ssoff = 0;                  // float offset of the tile-ring slot of output step j-2R

    for (int jb = 0; jb < NP; jb += Q) {
#pragma unroll
        for (int p = 0; p < Q; ++p) {
            const int j = jb + p;
            if (j >= NP) break;
            b2ptx::mbar_wait(&full[p], par);
            q[p] = b2ptx::lds128(my_col + p * PLANE);

            if (j >= 2 * R) {
                constexpr int dummy = 0;
                (void)dummy;
                const int pc = (p - R + Q) % Q;                // stage of the centre plane j-R
                const float *cp = my_col + pc * PLANE;
                const float4 c = q[pc];
                // z direction: row segment [-RZ, 4+RZ) around my 4 points
                float zr[4 + 2 * RZ];
#pragma unroll
                for (int m = 0; m < RZ / 4; ++m) {
                    // the outermost segments are only partly used when R is not a multiple of 4
                    // (R = 6: two floats each side): 8-byte loads halve their shared-memory wavefronts
                    if (R % 4 == 2 && m == 0) {
                        const float2 l = *reinterpret_cast<const float2 *>(cp - RZ + 2);
                        zr[0] = 0.f; zr[1] = 0.f; zr[2] = l.x; zr[3] = l.y;
                    } else {
                        const float4 l = b2ptx::lds128(cp - RZ + 4 * m);
                        zr[4 * m + 0] = l.x; zr[4 * m + 1] = l.y; zr[4 * m + 2] = l.z; zr[4 * m + 3] = l.w;
                    }
                    if (R % 4 == 2 && m == RZ / 4 - 1) {
                        const float2 r = *reinterpret_cast<const float2 *>(cp + 4 + 4 * m);
                        zr[RZ + 4 + 4 * m + 0] = r.x; zr[RZ + 4 + 4 * m + 1] = r.y;
                        zr[RZ + 4 + 4 * m + 2] = 0.f; zr[RZ + 4 + 4 * m + 3] = 0.f;
                    } else {
                        const float4 r = b2ptx::lds128(cp + 4 + 4 * m);
                        zr[RZ + 4 + 4 * m + 0] = r.x; zr[RZ + 4 + 4 * m + 1] = r.y;
                        zr[RZ + 4 + 4 * m + 2] = r.z; zr[RZ + 4 + 4 * m + 3] = r.w;
                    }
                }
                zr[RZ + 0] = c.x; zr[RZ + 1] = c.y; zr[RZ + 2] = c.z; zr[RZ + 3] = c.w;
                float4 acc = make_float4(wc * c.x, wc * c.y, wc * c.z, wc * c.w);
#pragma unroll
                for (int i = 1; i <= R; ++i) {
                    acc.x = fmaf(k.wz[i], zr[RZ + 0 - i] + zr[RZ + 0 + i], acc.x);
                    acc.y = fmaf(k.wz[i], zr[RZ + 1 - i] + zr[RZ + 1 + i], acc.y);
                    acc.z = fmaf(k.wz[i], zr[RZ + 2 - i] + zr[RZ + 2 + i], acc.z);
                    acc.w = fmaf(k.wz[i], zr[RZ + 3 - i] + zr[RZ + 3 + i], acc.w);
                }
#pragma unroll
                for (int i = 1; i <= R; ++i) {
                    const float4 a = b2ptx::lds128(cp - i * BZ);
                    const float4 bb = b2ptx::lds128(cp + i * BZ);
                    f4fma(acc, k.wy[i], f4add(a, bb));
                }
#pragma unroll
                for (int i = 1; i <= R; ++i)
                    f4fma(acc, k.wx[i], f4add(q[(pc - i + Q) % Q], q[(pc + i) % Q]));

                const float4 pv = b2ptx::lds128(my_prev + ssoff);
                const float4 ca = b2ptx::lds128(my_prev + (NS * TILE) + ssoff);
                const float4 dcp = f4sub(c, pv);
                float4 o;
                if (PK == B2_PARAM_SCALAR) {
                    o.x = fmaf(ca.x, fmaf(k.m_dt2, dcp.x, acc.x), c.x);
                    o.y = fmaf(ca.y, fmaf(k.m_dt2, dcp.y, acc.y), c.y);
                    o.z = fmaf(ca.z, fmaf(k.m_dt2, dcp.z, acc.z), c.z);
                    o.w = fmaf(ca.w, fmaf(k.m_dt2, dcp.w, acc.w), c.w);
                } else {
                    const float4 cb = b2ptx::lds128(my_prev + (2 * NS * TILE) + ssoff);
                    o.x = fmaf(ca.x, acc.x, fmaf(cb.x, dcp.x, c.x));
                    o.y = fmaf(ca.y, acc.y, fmaf(cb.y, dcp.y, c.y));
                    o.z = fmaf(ca.z, acc.z, fmaf(cb.z, dcp.z, c.z));
                    o.w = fmaf(ca.w, acc.w, fmaf(cb.w, dcp.w, c.w));
                }
                float *dst = outp + (long long)j * osx;
                if (zcnt == 4) {
                    *reinterpret_cast<float4 *>(dst) = o;
                } else if (zcnt > 0) {
                    dst[0] = o.x;
                    if (zcnt > 1) dst[1] = o.y;
                    if (zcnt > 2) dst[2] = o.z;
                }
                if (fuse) {
                    // a boundary plane: the same 16 bytes also go into the neighbour's halo (NVLink)
                    const int xo = xb + dx * (j - R);
                    float *pd = nullptr;
                    if (k.peer_lo && xo < k.pw) pd = k.peer_lo + (k.off_lo + (long long)xo * k.sx + rowoff);
                    if (k.peer_hi && xo >= k.nown - k.pw) pd = k.peer_hi + (k.off_hi + (long long)xo * k.sx + rowoff);
                    if (pd) {
                        if (zcnt == 4) {
                            *reinterpret_cast<float4 *>(pd) = o;
                        } else if (zcnt > 0) {
                            pd[0] = o.x;
                            if (zcnt > 1) pd[1] = o.y;
                            if (zcnt > 2) pd[2] = o.z;
                        }
                    }
                }
                ssoff += TILE;
                if (ssoff == NS * TILE) ssoff = 0;
            }
            if (j >= R) {
                __syncwarp();
                if (lane == 0) b2ptx::mbar_arrive(&empty[(p - R + Q) % Q]);
            }
        }
        par ^= 1;
    }
}


// ------------------------------------------------------------------------------------------
// k_iso_tma2: the same 2.5-D sweep with TWO y rows per consumer thread — for the large radii.
// ------------------------------------------------------------------------------------------
// k_iso_tma at so=12 (R=6) had to shrink its tile to 16x64 to fit the 13-stage plane ring plus the tile
// ring, and became shared-memory-bandwidth bound (ncu r1c: L1/TEX 80.6 %, HBM 75 %): 12 y-tap LDS.128 per
// float4 of output. Here
//   * a thread owns rows 2*tr and 2*tr+1: the 2R rows around them are loaded ONCE for both (7 LDS.128 per
//     output float4 instead of 12), each own row is the other's nearest y neighbour (registers);
//   * the tile is 28 x 64 (halo re-reads 1.79x instead of 2.19x): the plane ring holds R+1 live planes + 4 of
//     prefetch instead of 2R+1, u[t-1] and the coefficient tiles travel by TMA in a 5-stage tile ring (PF = 0,
//     the default: 3.08 ms at 1024^3 vs 3.20 ms for k_iso_tma). The first version read them straight from global
//     memory (LDG.128 prefetched two planes ahead, PF = 2): correct but 4.4 ms — with two warps per scheduler the
//     prefetched loads stall (ncu profiles/r2i_*: long-scoreboard) — kept as B2_ISO_V2=1 for the record;
//   * the x-history of the own columns is kept in register queues with static slots (planes c+1..c+R in
//     fut[(p+k) mod R], c-R..c-1 in pst[...], the loop unrolled R times — not 2R+1 = 13 times, which is
//     what made the first "taller tile" experiment of round 1 miss the instruction cache);
//   * y/x taps and the update in packed fp32x2 arithmetic.
// Halo step under decomposition: identical to k_iso_tma (peer stores, flag acquire, chunk 0 backwards).
template <int R, int TY, int TZ4, int PF = 2, int NT = 2>
struct IsoTma2Cfg {
    static constexpr int RZ = ceil4(R);
    static constexpr int TZ = 4 * TZ4;
    static constexpr int BY = TY + 2 * R;
    static constexpr int BZ = TZ + 2 * RZ;
    // PF == 0: u[t-1] and the coefficient tiles travel by TMA with the planes (like k_iso_tma): the plane ring is
    // R+1 live planes + 4 of prefetch, the tile ring 5 stages
    static constexpr int NU = PF == 0 ? R + 5 : 2 * R + 1;
    static constexpr int NS = PF == 0 ? 5 : 0;
    static constexpr int PLANE = align32f(BY * BZ);
    static constexpr int TILE = TY * TZ;
    static constexpr int NCT = (TY / 2) * TZ4;          // consumer threads
    static constexpr int NCW = NCT / 32;
    static constexpr size_t SMEM = (size_t)(NU * PLANE + NS * NT * TILE) * 4 + 2 * NU * 8 + 128;
};

template <int R, int TY, int TZ4, int PK, int PF>
__global__ void __launch_bounds__((TY / 2) * TZ4 + 32, 1)
k_iso_tma2(const __grid_constant__ CUtensorMap tm_uh, const __grid_constant__ CUtensorMap tm_uc,
           const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const IsoTK<R> k) {
    constexpr int NT = PK != B2_PARAM_SCALAR ? 3 : 2;
    using C = IsoTma2Cfg<R, TY, TZ4, PF, NT>;
    using b2ptx::F4;
    constexpr int RZ = C::RZ, TZ = C::TZ, BZ = C::BZ, NU = C::NU, PLANE = C::PLANE, NCW = C::NCW;
    constexpr int NS = C::NS, TILE = C::TILE;
    static_assert(TY % 2 == 0 && C::NCT % 32 == 0, "two rows per thread, whole warps");
    static_assert(R % 2 == 0, "the prefetch buffers alternate with the parity of the unrolled iteration");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_u = reinterpret_cast<float *>(smem_raw);
    float *s_t = s_u + NU * PLANE;                       // tile ring (PF == 0): [stage][prev, A, (B)][TILE]
    uint64_t *full = reinterpret_cast<uint64_t *>(s_t + NS * NT * TILE);
    uint64_t *empty = full + NU;

    int b = blockIdx.x;
    const int iz = b % k.ntz;
    b /= k.ntz;
    const int iy = b % k.nty;
    const int ix = b / k.nty;
    const int z0 = iz * TZ, y0 = iy * TY;
    const int xs = k.xlo + ix * k.lx;
    const int xe = min(xs + k.lx, k.xlo + k.xcount);
    const int NP = (xe - xs) + 2 * R;
    const bool back = (ix == 0) && k.back0;
    const int dx = back ? -1 : 1;
    const int xb = back ? xe - 1 + R : xs - R;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NU; ++i) {
            b2ptx::mbar_init(&full[i], 1);
            b2ptx::mbar_init(&empty[i], NCW);
        }
        b2ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCW) {
        if (lane == 0) {
            b2ptx::tma_prefetch_desc(&tm_uh);
            int slot = 0, round = 0, ss = 0;
            bool need_lo = k.want >= 0 && k.flag_lo != nullptr, need_hi = k.want >= 0 && k.flag_hi != nullptr;
            for (int j = 0; j < NP; ++j) {
                if (round > 0) b2ptx::mbar_wait(&empty[slot], (round - 1) & 1);
                const int xj = xb + dx * j;
                if (need_lo && xj < 0) {
                    while (b2ptx::ld_acquire_sys(k.flag_lo) < k.want) __nanosleep(40);
                    b2ptx::fence_proxy_async_global();
                    need_lo = false;
                }
                if (need_hi && xj >= k.nown) {
                    while (b2ptx::ld_acquire_sys(k.flag_hi) < k.want) __nanosleep(40);
                    b2ptx::fence_proxy_async_global();
                    need_hi = false;
                }
                const bool tiles = PF == 0 && j >= 2 * R;
                b2ptx::mbar_arrive_expect_tx(&full[slot], (uint32_t)(C::BY * BZ * 4) + (tiles ? (uint32_t)(TILE * 4 * NT) : 0u));
                b2ptx::tma_load_4d(s_u + slot * PLANE, &tm_uh, &full[slot], k.oz + z0 - RZ, k.oy + y0 - R,
                                   k.ox + xj, k.slot0);
                if (tiles) {
                    // the tiles of the output plane of iteration j travel with plane j
                    const int xo = xj - dx * R;
                    float *t = s_t + ss * NT * TILE;
                    b2ptx::tma_load_4d(t, &tm_uc, &full[slot], k.oz + z0, k.oy + y0, k.ox + xo, k.slotm);
                    b2ptx::tma_load_3d(t + TILE, &tm_a, &full[slot], k.oz + z0, k.oy + y0, k.ox + xo);
                    if (PK != B2_PARAM_SCALAR)
                        b2ptx::tma_load_3d(t + 2 * TILE, &tm_b, &full[slot], k.oz + z0, k.oy + y0, k.ox + xo);
                    if (++ss == NS) ss = 0;
                }
                if (++slot == NU) { slot = 0; ++round; }
            }
        }
        return;
    }

    // ---------------- consumers: rows (2 tr, 2 tr + 1), four z points ----------------
    const int tr = tid / TZ4, tz4 = tid % TZ4;
    const int gz = z0 + 4 * tz4;
    const int gy0 = y0 + 2 * tr;
    int zc[2];
    long long row[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        zc[r] = (gy0 + r < k.ny) ? min(max(k.nz - gz, 0), 4) : 0;
        row[r] = (long long)(k.oy + gy0 + r) * k.sy + (k.oz + gz);
    }
    const int col = (2 * tr + R) * BZ + RZ + 4 * tz4;      // row 0 of this thread inside a shared plane
    const long long osx = dx * k.sx;
    const long long gbase = (long long)(k.ox + xb - dx * R) * k.sx;   // + j * osx = the output plane of iteration j
    const bool fuse = (k.peer_lo != nullptr) | (k.peer_hi != nullptr);

    F4 fut[2][R], pst[2][R], cen[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        cen[r] = b2ptx::f4zero();
#pragma unroll
        for (int i = 0; i < R; ++i) { fut[r][i] = b2ptx::f4zero(); pst[r][i] = b2ptx::f4zero(); }
    }
    // u[t-1], A (and B) of the output plane of iteration j, loaded two iterations ahead
    constexpr int PFB = PF > 0 ? PF : 1;
    float4 pv[2][PFB], pa[2][PFB], pb[2][PFB];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int d = 0; d < PFB; ++d) { pv[r][d] = pa[r][d] = pb[r][d] = make_float4(0.f, 0.f, 0.f, 0.f); }
    int ts_off = 0;                                       // tile-ring stage of the current output plane (PF == 0)
    const int tcol = (2 * tr) * TZ + 4 * tz4;
    auto prefetch = [&](int j, int d) {
        // output plane of iteration j: x = xb + dx * (j - R), valid for 2R <= j < NP
        if (j < 2 * R || j >= NP) return;
        const long long g = gbase + (long long)j * osx;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (zc[r] > 0) {
                pv[r][d] = *reinterpret_cast<const float4 *>(k.um + g + row[r]);
                pa[r][d] = *reinterpret_cast<const float4 *>(k.cA + g + row[r]);
                if (PK != B2_PARAM_SCALAR) pb[r][d] = *reinterpret_cast<const float4 *>(k.cB + g + row[r]);
            }
        }
    };
    if (PF >= 1) prefetch(2 * R, 0);
    if (PF == 2) prefetch(2 * R + 1, PF - 1);

    int st_new = 0, st_cen = 0;           // shared-memory stage of plane j and of the centre plane j - R
    uint32_t par_new = 0;

    for (int jb = 0; jb < NP; jb += R) {
#pragma unroll
        for (int p = 0; p < R; ++p) {
            const int j = jb + p;
            if (j >= NP) break;
            // slot of plane c + kk (c = j - R, c mod R == p): compile-time
#define B2_SLOT(kk) ((((p) + (kk)) % R + R) % R)
            b2ptx::mbar_wait(&full[st_new], par_new);
            {
                const float *np_ = s_u + st_new * PLANE + col;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    pst[r][B2_SLOT(-1)] = cen[r];                        // plane c-1 (overwrites plane c-1-R)
                    cen[r] = fut[r][p];                                  // plane c
                    fut[r][p] = b2ptx::f4pack(b2ptx::lds128(np_ + r * BZ));   // plane c+R = j
                }
            }
            if (j >= 2 * R) {
                const float *cp0 = s_u + st_cen * PLANE + col;
                F4 acc[2];
                // z direction (scalar: the operands straddle register pairs)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float *cp = cp0 + r * BZ;
                    const float4 c = b2ptx::f4unpack(cen[r]);
                    float zr[4 + 2 * RZ];
#pragma unroll
                    for (int m = 0; m < RZ / 4; ++m) {
                        if (R % 4 == 2 && m == 0) {
                            const float2 l = *reinterpret_cast<const float2 *>(cp - RZ + 2);
                            zr[0] = 0.f; zr[1] = 0.f; zr[2] = l.x; zr[3] = l.y;
                        } else {
                            const float4 l = b2ptx::lds128(cp - RZ + 4 * m);
                            zr[4 * m + 0] = l.x; zr[4 * m + 1] = l.y; zr[4 * m + 2] = l.z; zr[4 * m + 3] = l.w;
                        }
                        if (R % 4 == 2 && m == RZ / 4 - 1) {
                            const float2 q = *reinterpret_cast<const float2 *>(cp + 4 + 4 * m);
                            zr[RZ + 4 + 4 * m + 0] = q.x; zr[RZ + 4 + 4 * m + 1] = q.y;
                            zr[RZ + 4 + 4 * m + 2] = 0.f; zr[RZ + 4 + 4 * m + 3] = 0.f;
                        } else {
                            const float4 q = b2ptx::lds128(cp + 4 + 4 * m);
                            zr[RZ + 4 + 4 * m + 0] = q.x; zr[RZ + 4 + 4 * m + 1] = q.y;
                            zr[RZ + 4 + 4 * m + 2] = q.z; zr[RZ + 4 + 4 * m + 3] = q.w;
                        }
                    }
                    zr[RZ + 0] = c.x; zr[RZ + 1] = c.y; zr[RZ + 2] = c.z; zr[RZ + 3] = c.w;
                    const float wc = k.p_wc.x;
                    float4 a = make_float4(wc * c.x, wc * c.y, wc * c.z, wc * c.w);
#pragma unroll
                    for (int i = 1; i <= R; ++i) {
                        a.x = fmaf(k.wz[i], zr[RZ + 0 - i] + zr[RZ + 0 + i], a.x);
                        a.y = fmaf(k.wz[i], zr[RZ + 1 - i] + zr[RZ + 1 + i], a.y);
                        a.z = fmaf(k.wz[i], zr[RZ + 2 - i] + zr[RZ + 2 + i], a.z);
                        a.w = fmaf(k.wz[i], zr[RZ + 3 - i] + zr[RZ + 3 + i], a.w);
                    }
                    acc[r] = b2ptx::f4pack(a);
                }
                // y direction: rows -R..-1 and +2..R+1 (relative to row 0) are loaded once for both rows;
                // each own row is the other's +-1 neighbour (registers)
                b2ptx::f4fma2(acc[0], k.p_wy[1], cen[1]);
                b2ptx::f4fma2(acc[1], k.p_wy[1], cen[0]);
#pragma unroll
                for (int d = -R; d <= R + 1; ++d) {
                    if (d == 0 || d == 1) continue;
                    const F4 v = b2ptx::f4pack(b2ptx::lds128(cp0 + d * BZ));
                    const int a0 = d < 0 ? -d : d, a1 = d - 1 < 0 ? 1 - d : d - 1;
                    if (a0 <= R) b2ptx::f4fma2(acc[0], k.p_wy[a0], v);
                    if (a1 <= R) b2ptx::f4fma2(acc[1], k.p_wy[a1], v);
                }
                // x direction: own columns, register queues
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int i = 1; i <= R; ++i)
                        b2ptx::f4fma2(acc[r], k.p_wx[i], b2ptx::f4add2(pst[r][B2_SLOT(-i)], fut[r][B2_SLOT(i)]));
                // update  u+ = u + A (m/dt^2 (u - u-) + lap)   |   u + B (u - u-) + A lap
                // compile-time buffer index (jb and 2R are even): a run-time index into a register array makes the
                // compiler select/copy right after the load, i.e. wait for it — the first version stalled on exactly that
                const int d = PF == 2 ? (p & 1) : 0;
                const long long g = gbase + (long long)j * osx;
                if (PF == 0) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float *t = s_t + ts_off + tcol + r * TZ;
                        pv[r][0] = b2ptx::lds128(t);
                        pa[r][0] = b2ptx::lds128(t + TILE);
                        if (PK != B2_PARAM_SCALAR) pb[r][0] = b2ptx::lds128(t + 2 * TILE);
                    }
                    ts_off += NT * TILE;
                    if (ts_off == NS * NT * TILE) ts_off = 0;
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const F4 c = cen[r];
                    const F4 dcp = b2ptx::f4sub2(c, b2ptx::f4pack(pv[r][d]));
                    const F4 ca = b2ptx::f4pack(pa[r][d]);
                    F4 o;
                    if (PK == B2_PARAM_SCALAR) {
                        F4 t = acc[r];
                        b2ptx::f4fma2(t, k.p_mdt2, dcp);
                        o = F4{b2ptx::fma2(ca.a, t.a, c.a), b2ptx::fma2(ca.b, t.b, c.b)};
                    } else {
                        const F4 cb = b2ptx::f4pack(pb[r][d]);
                        const F4 t = F4{b2ptx::fma2(cb.a, dcp.a, c.a), b2ptx::fma2(cb.b, dcp.b, c.b)};
                        o = F4{b2ptx::fma2(ca.a, acc[r].a, t.a), b2ptx::fma2(ca.b, acc[r].b, t.b)};
                    }
                    const float4 of = b2ptx::f4unpack(o);
                    float *dst = k.u1 + g + row[r];
                    if (zc[r] == 4) {
                        *reinterpret_cast<float4 *>(dst) = of;
                    } else if (zc[r] > 0) {
                        dst[0] = of.x;
                        if (zc[r] > 1) dst[1] = of.y;
                        if (zc[r] > 2) dst[2] = of.z;
                    }
                    if (fuse) {
                        const int xo = xb + dx * (j - R);
                        float *pd = nullptr;
                        const long long ro = row[r] - (long long)0;
                        if (k.peer_lo && xo < k.pw) pd = k.peer_lo + (k.off_lo + (long long)xo * k.sx + ro);
                        if (k.peer_hi && xo >= k.nown - k.pw) pd = k.peer_hi + (k.off_hi + (long long)xo * k.sx + ro);
                        if (pd) {
                            if (zc[r] == 4) {
                                *reinterpret_cast<float4 *>(pd) = of;
                            } else if (zc[r] > 0) {
                                pd[0] = of.x;
                                if (zc[r] > 1) pd[1] = of.y;
                                if (zc[r] > 2) pd[2] = of.z;
                            }
                        }
                    }
                }
                if (PF == 2) prefetch(j + 2, d);
            }
            if (PF == 1 && j + 1 >= 2 * R) prefetch(j + 1, 0);    // (after the stores of this plane were issued)
            if (j >= R) {
                __syncwarp();
                if (lane == 0) b2ptx::mbar_arrive(&empty[st_cen]);
                if (++st_cen == NU) st_cen = 0;
            }
            if (++st_new == NU) { st_new = 0; par_new ^= 1; }
#undef B2_SLOT
        }
    }
}

// A = 1/(m/dt^2 + damp/dt), B = m/dt^2 * A on the full allocated array (halo included; halo
// values are never used by the stencil)
__global__ void __launch_bounds__(256)
k_iso_coef(const float *__restrict__ damp, const float *__restrict__ param, int param_kind,
           float m_dt2_scalar, float inv_dt, float inv_dt2, float *__restrict__ A,
           float *__restrict__ B, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float md = m_dt2_scalar;
        if (param_kind == B2_PARAM_VP) {
            const float v = param[i];
            md = inv_dt2 / (v * v);
        } else if (param_kind == B2_PARAM_M) {
            md = param[i] * inv_dt2;
        }
        const float a = 1.0f / (md + damp[i] * inv_dt);
        A[i] = a;
        if (B) B[i] = md * a;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

// tensor map over a (t?, x, y, z) f32 array; box = (bz, by, 1[, 1])
static int make_tmap(CUtensorMap *tm, const void *base, int rank, const int *dims_zyxt,
                     int bz, int by) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return B2_ERR_DEVICE; }
    cuuint64_t gdim[4];
    cuuint64_t gstr[3];
    cuuint32_t box[4] = {(cuuint32_t)bz, (cuuint32_t)by, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cuuint64_t stride = 4;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = (cuuint64_t)dims_zyxt[i];
        stride *= gdim[i];
        if (i < rank - 1) gstr[i] = stride;
    }
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base),
                     gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rank=%d box=(%d,%d)", (int)r, rank, bz, by);
        return B2_ERR_DEVICE;
    }
    return B2_OK;
}

// tile configuration per radius (TY, TZ4): sized so that the Q-stage plane ring plus the
// (R+1)-stage tile ring fit the 227 KB of shared memory of one SM
template <int R> struct TileOf;
template <> struct TileOf<2> { static constexpr int TY = 32, TZ4 = 16; };
template <> struct TileOf<4> { static constexpr int TY = 32, TZ4 = 16; };
template <> struct TileOf<6> { static constexpr int TY = 16, TZ4 = 16; };
template <> struct TileOf<8> { static constexpr int TY = 8, TZ4 = 16; };

// k_iso_tma2 (two rows per thread, no tile ring): the radii whose k_iso_tma tile had to shrink
template <int R> struct Tile2Of { static constexpr bool on = false; static constexpr int TY = 32, TZ4 = 16; };
// 28 x 64: 14 x 16 = 224 consumer threads + the producer warp = 8 warps, so that the register file splits
// into 255 registers per thread (9 warps would be allocated as 12: 168 registers, and the queues spill)
template <> struct Tile2Of<6> { static constexpr bool on = true; static constexpr int TY = 28, TZ4 = 16; };


static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// scratch for the coefficient arrays, cached across calls (re-allocated when the size changes)
static float *g_coef[3] = {nullptr, nullptr, nullptr};     // A, B (TMA path), W (OT4)
static size_t g_coef_elems[3] = {0, 0, 0};

static int coef_buffer(int which, size_t elems, float **out) {
    if (g_coef_elems[which] != elems) {
        if (g_coef[which]) cudaFree(g_coef[which]);
        g_coef[which] = nullptr;
        g_coef_elems[which] = 0;
        B2_CUDA(cudaMalloc(&g_coef[which], elems * sizeof(float)), B2_ERR_MEMORY);
        g_coef_elems[which] = elems;
    }
    *out = g_coef[which];
    return B2_OK;
}

// rows [y0, y1) of the planes [x0, x1) (allocated indices): the y-skewed streamed loop tabulates row blocks
__global__ void __launch_bounds__(256)
k_iso_coef_rows(const float *__restrict__ damp, const float *__restrict__ param, int param_kind, float m_dt2_scalar,
                float inv_dt, float inv_dt2, float *__restrict__ A, float *__restrict__ B, long long sx, int a2,
                int x0, int y0, int ny) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= a2) return;
    const int x = x0 + blockIdx.z;
    for (int y = y0 + blockIdx.y; y < y0 + ny; y += gridDim.y) {
        const size_t i = (size_t)x * (size_t)sx + (size_t)y * a2 + z;
        float md = m_dt2_scalar;
        if (param_kind == B2_PARAM_VP) {
            const float v = param[i];
            md = inv_dt2 / (v * v);
        } else if (param_kind == B2_PARAM_M) {
            md = param[i] * inv_dt2;
        }
        const float a = 1.0f / (md + damp[i] * inv_dt);
        A[i] = a;
        if (B) B[i] = md * a;
    }
}

int iso_coef_tabulate_rows(const IsoPlan &p, int x0, int x1, int y0, int y1) {
    if (x1 <= x0 || y1 <= y0) return B2_OK;
    const float inv_dt = 1.0f / p.dt, inv_dt2 = 1.0f / (p.dt * p.dt);
    const float md = (1.0f / (p.vp * p.vp)) * inv_dt2;
    dim3 block(256, 1, 1), grid((p.a[2] + 255) / 256, (unsigned)std::min(y1 - y0, 1024), (unsigned)(x1 - x0));
    k_iso_coef_rows<<<grid, block, 0, stream()>>>(p.damp, p.param, p.param_kind, md, inv_dt, inv_dt2, p.coefA, p.coefB,
                                                  p.sx, p.a[2], x0, y0, y1 - y0);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

// tabulate the allocated x-planes [plane_lo, plane_hi) (the streamed time loop does it chunk by chunk, as
// the damping / parameter planes arrive from the host)
int iso_coef_tabulate_planes(const IsoPlan &p, int plane_lo, int plane_hi) {
    if (plane_hi <= plane_lo) return B2_OK;
    const float inv_dt = 1.0f / p.dt, inv_dt2 = 1.0f / (p.dt * p.dt);
    const float md = (1.0f / (p.vp * p.vp)) * inv_dt2;
    const size_t off = (size_t)plane_lo * (size_t)p.sx, cnt = (size_t)(plane_hi - plane_lo) * (size_t)p.sx;
    k_iso_coef<<<148 * 8, 256, 0, stream()>>>(p.damp + off, p.param ? p.param + off : nullptr, p.param_kind, md, inv_dt,
                                               inv_dt2, p.coefA + off, p.coefB ? p.coefB + off : nullptr, cnt);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

static int iso_coef_tabulate(IsoPlan &p) {
    int rc;
    if ((rc = coef_buffer(0, p.slot_elems, &p.coefA))) return rc;
    p.coefB = nullptr;
    if (p.param_kind != B2_PARAM_SCALAR)
        if ((rc = coef_buffer(1, p.slot_elems, &p.coefB))) return rc;
    if (p.defer_coef) return B2_OK;          // the caller tabulates plane ranges itself
    return iso_coef_tabulate_planes(p, 0, p.a[0]);
}

template <int R>
static int plan_tma(IsoPlan &p) {
    using T = TileOf<R>;
    constexpr int RZ = ceil4(R);
    const int dims4[4] = {p.a[2], p.a[1], p.a[0], p.tsize};
    const int dims3[3] = {p.a[2], p.a[1], p.a[0]};
    int rc;
    p.v2 = Tile2Of<R>::on && p.n[1] >= 16 ? env_int("B2_ISO_V2", 4) : 0;   // so=12: variant 4 (3.08 vs 3.20 ms at 1024^3)
    if (p.v2) {
        using T2 = Tile2Of<R>;
        if ((rc = make_tmap(&p.tm_uh, p.u, 4, dims4, 4 * T2::TZ4 + 2 * RZ, T2::TY + 2 * R))) return rc;
        // variant 4: u[t-1] and the coefficient tiles by TMA (28 x 64 tiles)
        if ((rc = make_tmap(&p.tm_uc, p.u, 4, dims4, 4 * T2::TZ4, T2::TY))) return rc;
        if ((rc = make_tmap(&p.tm_damp, p.coefA, 3, dims3, 4 * T2::TZ4, T2::TY))) return rc;
        if (p.param_kind != B2_PARAM_SCALAR) {
            if ((rc = make_tmap(&p.tm_par, p.coefB, 3, dims3, 4 * T2::TZ4, T2::TY))) return rc;
        } else {
            p.tm_par = p.tm_damp;
        }
        return B2_OK;
    }
    if ((rc = make_tmap(&p.tm_uh, p.u, 4, dims4, 4 * T::TZ4 + 2 * RZ, T::TY + 2 * R))) return rc;
    if ((rc = make_tmap(&p.tm_uc, p.u, 4, dims4, 4 * T::TZ4, T::TY))) return rc;
    if ((rc = make_tmap(&p.tm_damp, p.coefA, 3, dims3, 4 * T::TZ4, T::TY))) return rc;
    if (p.param_kind != B2_PARAM_SCALAR) {
        if ((rc = make_tmap(&p.tm_par, p.coefB, 3, dims3, 4 * T::TZ4, T::TY))) return rc;
    } else {
        p.tm_par = p.tm_damp;
    }
    return B2_OK;
}

int iso_plan_init(IsoPlan &p, int kernel) {
    const int R = p.radius[2];
    bool ok = (p.radius[0] == R && p.radius[1] == R) && (R == 2 || R == 4 || R == 6 || R == 8);
    ok = ok && p.damp != nullptr;
    ok = ok && (p.a[2] % 4 == 0) && (p.o[2] % 4 == 0);
    ok = ok && ((uintptr_t)p.u % 16 == 0) && ((uintptr_t)p.damp % 16 == 0);
    ok = ok && (p.param_kind == B2_PARAM_SCALAR || (uintptr_t)p.param % 16 == 0);
    ok = ok && (p.slot_elems % 4 == 0);
    // tiny grids gain nothing from the pipelined kernel
    ok = ok && (p.n[1] >= 8 && p.n[2] >= 16);
    if (kernel == 1) ok = false;
    if (p.ot4) {
        // two-pass generic path: W = lap(u)/m over the box grown by R, then the update with lap(W)
        if (kernel == 2) { set_error("iso: the OT4 kernel has no TMA variant yet"); return B2_ERR_INVALID; }
        for (int d = 0; d < 3; ++d) {
            const int rd = p.radius[d];
            if (p.o[d] - 2 * rd < 0 || p.o[d] + p.n[d] - 1 + 2 * rd >= p.a[d]) {
                set_error("iso: OT4 needs a halo of 2*radius = %d points on dim %d", 2 * rd, d);
                return B2_ERR_INVALID;
            }
        }
        p.use_tma = false;
        return coef_buffer(2, p.slot_elems, &p.ot4W);
    }
    if (kernel == 2 && !ok) {
        set_error("iso: TMA kernel forced but layout does not qualify (radius=%d a2=%d o2=%d)", R,
                  p.a[2], p.o[2]);
        return B2_ERR_INVALID;
    }
    p.use_tma = ok;
    if (!ok) return B2_OK;
    int rc = iso_coef_tabulate(p);
    if (rc) return rc;
    switch (R) {
        case 2: return plan_tma<2>(p);
        case 4: return plan_tma<4>(p);
        case 6: return plan_tma<6>(p);
        case 8: return plan_tma<8>(p);
    }
    return B2_ERR_INVALID;
}

template <int R, int PK>
static int launch_tma(const IsoPlan &p, int slot0, int slotm, int slot1, int xlo, int xcount,
                      const IsoFuse *fz = nullptr) {
    using T = TileOf<R>;
    using T2 = Tile2Of<R>;
    using C = IsoTmaCfg<R, T::TY, T::TZ4, PK>;
    using C2 = IsoTma2Cfg<R, T2::TY, T2::TZ4>;
    auto kern = k_iso_tma<R, T::TY, T::TZ4, PK>;
    static bool attr_set = false;
    if (!attr_set) {
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM),
                B2_ERR_LAUNCH);
        if constexpr (T2::on) {
            B2_CUDA(cudaFuncSetAttribute(k_iso_tma2<R, T2::TY, T2::TZ4, PK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C2::SMEM), B2_ERR_LAUNCH);
            using C2t = IsoTma2Cfg<R, T2::TY, T2::TZ4, 0, (PK != B2_PARAM_SCALAR ? 3 : 2)>;
            B2_CUDA(cudaFuncSetAttribute(k_iso_tma2<R, T2::TY, T2::TZ4, PK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C2t::SMEM), B2_ERR_LAUNCH);
        }
        attr_set = true;
    }
    const int v2 = T2::on ? p.v2 : 0;           // 4: TMA tile ring (default at so=12), 1: LDG prefetch (experiment)
    const int TYeff = v2 ? T2::TY : T::TY;
    IsoTK<R> k;
    k.u1 = p.u + (size_t)slot1 * p.slot_elems;
    k.sx = p.sx;
    k.sy = p.sy;
    k.ny = p.n[1];
    k.nz = p.n[2];
    k.ox = p.o[0];
    k.oy = p.o[1];
    k.oz = p.o[2];
    k.xlo = xlo;
    k.xcount = xcount;
    k.ntz = (p.n[2] + C::TZ - 1) / C::TZ;
    k.nty = (p.n[1] + TYeff - 1) / TYeff;
    // x-chunk length: enough CTAs for ~16 waves of 148 SMs, but chunks of at least 32 planes
    int lx = env_int("B2_ISO_LX", 0);
    if (lx <= 0) lx = choose_chunk_len(k.ntz * k.nty, xcount, 2 * R, 32);
    lx = std::min(lx, xcount);
    // fused halo step with neighbours on both sides: at least two chunks, so that the one marching down
    // to x = 0 and the one marching up to x = n-1 each meet the neighbour's planes at their end
    if (fz && (fz->peer_lo || fz->flag_lo) && (fz->peer_hi || fz->flag_hi) && lx >= xcount && xcount >= 4 * R)
        lx = (xcount + 1) / 2;
    k.lx = lx;
    const int ntx = (xcount + lx - 1) / lx;
    k.back0 = 0; k.nown = p.n[0]; k.pw = R;
    k.peer_lo = k.peer_hi = nullptr;
    k.off_lo = k.off_hi = 0;
    k.flag_lo = k.flag_hi = nullptr;
    k.want = -1;
    if (fz) {
        const long long plane = p.sx;
        k.back0 = (fz->peer_lo || fz->flag_lo) ? 1 : 0;
        k.peer_lo = fz->peer_lo;
        k.peer_hi = fz->peer_hi;
        // my owned plane x mirrors plane (halo + n_lo + x) of the low neighbour / (halo - nown + x) of the high one
        k.off_lo = (long long)slot1 * fz->slot_lo + (long long)(p.o[0] + fz->n_lo) * plane;
        k.off_hi = (long long)slot1 * fz->slot_hi + (long long)(p.o[0] - p.n[0]) * plane;
        k.flag_lo = fz->flag_lo;
        k.flag_hi = fz->flag_hi;
        k.want = fz->want;
    }
    k.slot0 = slot0;
    k.slotm = slotm;
    k.m_dt2 = (1.0f / (p.vp * p.vp)) * (1.0f / (p.dt * p.dt));
    for (int i = 0; i <= R; ++i) {
        k.wx[i] = p.w[0][i];
        k.wy[i] = p.w[1][i];
        k.wz[i] = p.w[2][i];
        k.p_wx[i] = make_float2(k.wx[i], k.wx[i]);
        k.p_wy[i] = make_float2(k.wy[i], k.wy[i]);
    }
    k.p_mdt2 = make_float2(k.m_dt2, k.m_dt2);
    {
        const float wc = k.wx[0] + k.wy[0] + k.wz[0];
        k.p_wc = make_float2(wc, wc);
    }
    k.um = p.u + (size_t)slotm * p.slot_elems;
    k.cA = p.coefA;
    k.cB = p.coefB;
    const unsigned grid = (unsigned)(k.ntz * k.nty * ntx);
    timing_begin();
    bool launched = false;
    if constexpr (T2::on) {
        using C2t = IsoTma2Cfg<R, T2::TY, T2::TZ4, 0, (PK != B2_PARAM_SCALAR ? 3 : 2)>;
        const unsigned nth = (T2::TY / 2) * T2::TZ4 + 32;
        if (v2 == 1) k_iso_tma2<R, T2::TY, T2::TZ4, PK, 2><<<grid, nth, C2::SMEM, stream()>>>(p.tm_uh, p.tm_uc, p.tm_damp, p.tm_par, k);
        if (v2 == 4) k_iso_tma2<R, T2::TY, T2::TZ4, PK, 0><<<grid, nth, C2t::SMEM, stream()>>>(p.tm_uh, p.tm_uc, p.tm_damp, p.tm_par, k);
        launched = v2 == 1 || v2 == 4;
    }
    if (!launched)
        kern<<<grid, T::TY * T::TZ4 + 32, C::SMEM, stream()>>>(p.tm_uh, p.tm_uc, p.tm_damp, p.tm_par, k);
    timing_end();
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

template <int R>
static int launch_tma_pk(const IsoPlan &p, int s0, int sm, int s1, int xlo, int xcount, const IsoFuse *fz = nullptr) {
    switch (p.param_kind) {
        case B2_PARAM_SCALAR: return launch_tma<R, B2_PARAM_SCALAR>(p, s0, sm, s1, xlo, xcount, fz);
        case B2_PARAM_VP: return launch_tma<R, B2_PARAM_VP>(p, s0, sm, s1, xlo, xcount, fz);
        case B2_PARAM_M: return launch_tma<R, B2_PARAM_M>(p, s0, sm, s1, xlo, xcount, fz);
    }
    return B2_ERR_INVALID;
}

static IsoGK generic_args(const IsoPlan &p, int slot0, int slotm, int slot1, int xlo, int xcount) {
    IsoGK k;
    k.u0 = p.u + (size_t)slot0 * p.slot_elems;
    k.um = p.u + (size_t)slotm * p.slot_elems;
    k.u1 = p.u + (size_t)slot1 * p.slot_elems;
    k.damp = p.damp;
    k.param = p.param;
    k.sx = p.sx;
    k.sy = p.sy;
    k.n0 = xcount;
    k.n1 = p.n[1];
    k.n2 = p.n[2];
    k.o0 = p.o[0] + xlo;
    k.o1 = p.o[1];
    k.o2 = p.o[2];
    k.r0 = p.radius[0];
    k.r1 = p.radius[1];
    k.r2 = p.radius[2];
    k.param_kind = p.param_kind;
    k.inv_dt = 1.0f / p.dt;
    k.inv_dt2 = 1.0f / (p.dt * p.dt);
    k.m_dt2 = (1.0f / (p.vp * p.vp)) * k.inv_dt2;
    memcpy(k.w, p.w, sizeof(k.w));
    k.W = nullptr;
    k.ot4c = p.dt * p.dt / 12.0f;
    k.vp2 = p.vp * p.vp;
    k.U1 = nullptr;
    k.dm = nullptr;
    k.dsx = k.dsy = 0;
    k.dg0 = k.dg1 = k.dg2 = 0;
    return k;
}

int iso_born_source(const IsoPlan &p, int slot0, int slotm, int slot1, float *U1, const float *dm,
                    long long dsx, long long dsy, int dg0, int dg1, int dg2) {
    IsoGK k = generic_args(p, slot0, slotm, slot1, 0, p.n[0]);
    k.U1 = U1;
    k.dm = dm;
    k.dsx = dsx; k.dsy = dsy;
    k.dg0 = dg0; k.dg1 = dg1; k.dg2 = dg2;
    dim3 block(64, 4, 1);
    dim3 grid((k.n2 + 63) / 64, (k.n1 + 3) / 4, (unsigned)std::min(k.n0, 65535));
    k_born_src<<<grid, block, 0, stream()>>>(k);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

int iso_fs_fix(const IsoPlan &p, int slot0, int slotm, int slot1, int xlo, int xcount) {
    if (xcount <= 0) return B2_OK;
    if (p.n[2] <= 2 * p.radius[2]) {
        set_error("free surface: the vertical extent %d is too small for radius %d", p.n[2], p.radius[2]);
        return B2_ERR_INVALID;
    }
    IsoGK k = generic_args(p, slot0, slotm, slot1, xlo, xcount);
    dim3 block(16, 16, 1);
    dim3 grid((k.n1 + 15) / 16, (unsigned)std::min(xcount, 65535), 1);
    k_iso_fs_fix<<<grid, block, 0, stream()>>>(k);
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

int iso_step_fused(const IsoPlan &p, int slot0, int slotm, int slot1, const IsoFuse &f) {
    if (!p.use_tma) { set_error("iso: the fused halo step needs the TMA sweep kernel"); return B2_ERR_INVALID; }
    switch (p.radius[2]) {
        case 2: return launch_tma_pk<2>(p, slot0, slotm, slot1, 0, p.n[0], &f);
        case 4: return launch_tma_pk<4>(p, slot0, slotm, slot1, 0, p.n[0], &f);
        case 6: return launch_tma_pk<6>(p, slot0, slotm, slot1, 0, p.n[0], &f);
        case 8: return launch_tma_pk<8>(p, slot0, slotm, slot1, 0, p.n[0], &f);
    }
    return B2_ERR_INVALID;
}

int iso_step(const IsoPlan &p, int slot0, int slotm, int slot1, int xlo, int xcount) {
    if (xcount <= 0) return B2_OK;
    if (p.use_tma) {
        switch (p.radius[2]) {
            case 2: return launch_tma_pk<2>(p, slot0, slotm, slot1, xlo, xcount);
            case 4: return launch_tma_pk<4>(p, slot0, slotm, slot1, xlo, xcount);
            case 6: return launch_tma_pk<6>(p, slot0, slotm, slot1, xlo, xcount);
            case 8: return launch_tma_pk<8>(p, slot0, slotm, slot1, xlo, xcount);
        }
        return B2_ERR_INVALID;
    }
    IsoGK k = generic_args(p, slot0, slotm, slot1, xlo, xcount);
    if (p.ot4) {
        IsoGK g = k;                       // first pass on the box grown by the radius
        g.W = p.ot4W;
        g.o0 -= g.r0; g.o1 -= g.r1; g.o2 -= g.r2;
        g.n0 += 2 * g.r0; g.n1 += 2 * g.r1; g.n2 += 2 * g.r2;
        dim3 blk(64, 4, 1);
        dim3 grd((g.n2 + 63) / 64, (g.n1 + 3) / 4, (unsigned)std::min(g.n0, 65535));
        k_ot4_w<<<grd, blk, 0, stream()>>>(g);
        count_launch();
        B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
        k.W = p.ot4W;
    }
    dim3 block(64, 4, 1);
    dim3 grid((k.n2 + 63) / 64, (k.n1 + 3) / 4, (unsigned)std::min(xcount, 65535));
    timing_begin();
    k_iso_generic<<<grid, block, 0, stream()>>>(k);
    timing_end();
    count_launch();
    B2_CUDA(cudaGetLastError(), B2_ERR_LAUNCH);
    return B2_OK;
}

}  // namespace b2
