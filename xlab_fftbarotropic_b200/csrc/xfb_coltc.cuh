// xfb_coltc.cuh -- K-COL of the stepper at NX = 8192: thread-block CLUSTER of two CTAs per two-column tile.
//
// A two-column tile (whole 32-byte sectors in the pair layout, see xfb_colt.cuh) is 131 KB; together with the
// Stockham buffer of one column it fills an SM, which forced colt_kernel<8192> to use ONE staging buffer for both
// directions (the next tile's fetch waits for the last store) and to transform the two columns one after the other
// (the stage state of the other column has to be re-read for every product).  Here the tile is shared by the two
// CTAs of a cluster (two SMs):
//   * CTA r transforms column r only: one butterfly per thread per pass, the new stage state stays in registers
//     for the four products (no re-reads);
//   * CTA r moves the row-pair half r of the tile with TMA, all columns: it fetches [2048 pairs][2 cols][2 rows]
//     (64 KB) into SI and stores the same shape from SO -- whole sectors on both sides;
//   * the column <-> row-half redistribution goes through distributed shared memory: a thread reads/writes the 8
//     of its 16 rows that live in the other CTA's half directly in the peer's SI / SO (ld/st.shared::cluster).
//   * SI is free as soon as both CTAs have read it, so the next tile is fetched during the whole current tile.
// Cluster barriers order the DSMEM traffic (two per tile + two per product).
//
// STATUS: correct (same results as colt_kernel), but measured slower on B200 -- 1.14 ms against 0.92 ms per launch at
// 8192^2 -- so it is opt-in (XFB_COL_CLUSTER=1).  40 eight-byte DSMEM accesses per thread per tile and ten cluster
// barriers cost more than the state re-reads and the exposed tile fetch they remove.
#pragma once
#include <cooperative_groups.h>

#include "xfb_colt.cuh"

namespace xfb {

namespace cg = cooperative_groups;

template <int NX>
struct ColTCCfg {
    static constexpr int G = NX / 16;
    static constexpr int THREADS = G;                                    // one column per CTA
    static constexpr int HPAIRS = NX / 4;                                // row pairs per half
    static constexpr int BOXR = 256;
    static constexpr int NBOX = HPAIRS / BOXR;                           // TMA boxes per half tile
    static constexpr int H_BYTES = HPAIRS * 4 * (int)sizeof(cpx);        // [pairs][2 cols][2 rows]
    static constexpr int F_BYTES = LinePlan<NX>::PADDED * (int)sizeof(cpx);
    static constexpr int SMEM = 2 * H_BYTES + F_BYTES + 1024;
    static_assert(THREADS == 512 && HPAIRS % BOXR == 0, "cluster kernel is laid out for NX = 8192");
};

template <int NX, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ColTCCfg<NX>::THREADS, 1)
coltc_kernel(const ColParams p, const __grid_constant__ ColTMaps maps, const int tiles_per_member, const int tiles_total)
{
    typedef ColTCCfg<NX> C;
    constexpr int G = C::G;
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();          // my column of the tile and my row-pair half
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024 - (smem_u32(smem_dyn) & 1023)) & 1023);
    cpx *SI = reinterpret_cast<cpx *>(smem_raw);                                  // incoming half tile
    cpx *SO = reinterpret_cast<cpx *>(smem_raw + C::H_BYTES);                     // outgoing half tile
    cpx *F = reinterpret_cast<cpx *>(smem_raw + 2 * C::H_BYTES);                  // Stockham buffer, one column
    cpx *SI_peer = cluster.map_shared_rank(SI, r ^ 1);
    cpx *SO_peer = cluster.map_shared_rank(SO, r ^ 1);
    __shared__ unsigned long long full;

    const int tid = threadIdx.x;
    int t[1] = {tid}, c[1] = {0};
    LineTw<NX> tw[1];
    tw[0].init(p.tw, p.twn, tid);
    if (tid == 0) {
        mbar_init(&full, 1);
        mbar_fence_init();
    }
    cluster.sync();
    unsigned phase = 0;

    // rows of this thread: i = tid + k*G; pair = (tid >> 1) + k*256; k < 8 -> half 0, k >= 8 -> half 1
    // element (local pair lp, col, row-in-pair) of a half buffer: lp*4 + 2*col + (i & 1)
    const int h_base = (tid >> 1) * 4 + 2 * r + (tid & 1);
    const cpx *in_lo = (r == 0) ? SI : SI_peer, *in_hi = (r == 0) ? SI_peer : SI;      // halves 0 / 1 of the incoming tile
    cpx *out_lo = (r == 0) ? SO : SO_peer, *out_hi = (r == 0) ? SO_peer : SO;

    const int ncl = gridDim.x >> 1;
    int tile = blockIdx.x >> 1;
    if (MODE == COL_STEP && tile < tiles_total && tid == 0) {
        const int member = tile / tiles_per_member, tl = tile - member * tiles_per_member;
        mbar_expect_tx(&full, C::H_BYTES);
#pragma unroll 1
        for (int b = 0; b < C::NBOX; ++b)
            tma_load_2d(SI + (size_t)b * C::BOXR * 4, &maps.jint, tl * 4, member * (NX / 2) + r * C::HPAIRS + b * C::BOXR, &full);
    }

    for (; tile < tiles_total; tile += ncl) {
        const int member = tile / tiles_per_member, tl = tile - member * tiles_per_member;
        const size_t moff = (size_t)member * (size_t)p.member_stride;
        const int tmy = member * (NX / 2) + r * C::HPAIRS, tmx = tl * 4;
        const int j = p.j_base + tl * 2 + r;
        const float kyv = __ldg(p.ky + j);
        const float ky2 = kyv * kyv;
        // state arrays are tile-major with tile width 1: column (tl*2 + r) is one contiguous block of NX values
        const size_t e0 = moff + (size_t)(tl * 2 + r) * (size_t)p.st_tile_stride + (size_t)tid;
        cpx v[1][16];
        cpx zkeep[16];

        if (MODE == COL_STEP) {
            {
                const int bytes = NX * (int)sizeof(cpx);
                for (int o = tid * 128; o < bytes; o += C::THREADS * 128) {
                    prefetch_l2(reinterpret_cast<const char *>(p.z0 + e0 - tid) + o);
                    if (p.stage != 1) {
                        prefetch_l2(reinterpret_cast<const char *>(p.zk + e0 - tid) + o);
                        prefetch_l2(reinterpret_cast<const char *>(p.acc + e0 - tid) + o);
                    }
                }
            }
            mbar_wait(&full, phase);
            phase ^= 1;
            cluster.sync();                                   // both halves of the tile have landed
#pragma unroll
            for (int k = 0; k < 8; ++k) v[0][k] = in_lo[h_base + k * (256 * 4)];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[0][8 + k] = in_hi[h_base + k * (256 * 4)];
            cluster.sync();                                   // both CTAs have read both halves: SI is free
            {
                const int nt = tile + ncl;
                if (nt < tiles_total && tid == 0) {
                    const int nm = nt / tiles_per_member, ntl = nt - nm * tiles_per_member;
                    mbar_expect_tx(&full, C::H_BYTES);
#pragma unroll 1
                    for (int b = 0; b < C::NBOX; ++b)
                        tma_load_2d(SI + (size_t)b * C::BOXR * 4, &maps.jint, ntl * 4, nm * (NX / 2) + r * C::HPAIRS + b * C::BOXR, &full);
                }
            }
            col_fft<NX, 1, 1>(v, F, t, c, tw);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                cpx z0v[8], zkv[8], av[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) z0v[q] = p.z0[e0 + (size_t)((8 * h + q) * G)];
                if (p.stage != 1) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        zkv[q] = p.zk[e0 + (size_t)((8 * h + q) * G)];
                        av[q] = p.acc[e0 + (size_t)((8 * h + q) * G)];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) { zkv[q] = z0v[q]; av[q] = mk(0.f, 0.f); }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = 8 * h + q;
                    const int i = tid + k * G;
                    const size_t e = e0 + (size_t)(k * G);
                    const cpx X = v[0][k];
                    const float kxv = (float)signed_row<NX>(tid, k) * p.kxscale;
                    const float lap = -fmaf(kxv, kxv, ky2);
                    // dvortdt_c += (vort_c * laplacian_coe) * NU                                  main.cpp:240-243
                    const float tx = __fadd_rn(X.x, __fmul_rn(__fmul_rn(zkv[q].x, lap), p.nu));
                    const float ty2 = __fadd_rn(X.y, __fmul_rn(__fmul_rn(zkv[q].y, lap), p.nu));
                    // dealiasing mask (fftwfop.cpp:57-68)
                    const int ii = (i <= NX / 2) ? i : NX - i;
                    const float m = (ii * ii + j * j >= p.mask_kd_i) ? 0.0f : 1.0f;
                    const float rx = __fmul_rn(tx, m), ry = __fmul_rn(ty2, m);
                    cpx zn;
                    if (p.stage == 4) {                                                          // main.cpp:309-312
                        zn.x = __fadd_rn(z0v[q].x, __fdiv_rn(__fmul_rn(__fadd_rn(av[q].x, rx), p.dt), 6.0f));
                        zn.y = __fadd_rn(z0v[q].y, __fdiv_rn(__fmul_rn(__fadd_rn(av[q].y, ry), p.dt), 6.0f));
                        p.z0[e] = zn;
                    } else {                                                                     // main.cpp:246-251
                        const cpx an = (p.stage == 1) ? mk(rx, ry)
                                                      : mk(__fadd_rn(av[q].x, __fmul_rn(2.0f, rx)),
                                                           __fadd_rn(av[q].y, __fmul_rn(2.0f, ry)));
                        p.acc[e] = an;
                        zn.x = __fadd_rn(z0v[q].x, __fmul_rn(rx, p.dt_stage));
                        zn.y = __fadd_rn(z0v[q].y, __fmul_rn(ry, p.dt_stage));
                        p.zk[e] = zn;
                    }
                    zkeep[k] = zn;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) zkeep[k] = p.z0[e0 + (size_t)(k * G)];
        }

        // ------------------------------------------------------------------ prologue of the next stage + 4 inverse
#pragma unroll 1
        for (int f = 0; f < 4; ++f) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int i = tid + k * G;
                const cpx z = zkeep[k];
                const float kx = (float)signed_row<NX>(tid, k) * p.kxscale;
                // f = 0: i kx Z, 1: i ky Z, 2: i ky Psi (u before negation), 3: i kx Psi (v);
                // Psi = Z / -(kx^2+ky^2), (0,0) entry divides by 1                 (fftwfop.cpp:43,112-117)
                float kk = (f == 0 || f == 3) ? kx : kyv;
                if (f >= 2) {
                    const float li = (i == 0 && j == 0) ? 1.0f : -fmaf(kx, kx, ky2);
                    kk = __fdividef(kk, li);
                }
                v[0][k] = mk(z.x * kk, -z.y * kk);       // swap(i kk z)
            }
            col_fft<NX, 1, 1>(v, F, t, c, tw);
            if (tid == 0) tma_wait_read_all();                // my SO has been read by the previous store (long ago)
            cluster.sync();                                   // both SO buffers are free
#pragma unroll
            for (int k = 0; k < 8; ++k) out_lo[h_base + k * (256 * 4)] = cswap(v[0][k]);
#pragma unroll
            for (int k = 0; k < 8; ++k) out_hi[h_base + k * (256 * 4)] = cswap(v[0][8 + k]);
            asm volatile("fence.proxy.async;" ::: "memory");  // generic writes (local and remote) before the TMA reads
            cluster.sync();                                   // every write into my SO has landed
            if (tid == 0) {
                const CUtensorMap *mt = &maps.t[f];
#pragma unroll 1
                for (int b = 0; b < C::NBOX; ++b) tma_store_2d(mt, tmx, tmy + b * C::BOXR, SO + (size_t)b * C::BOXR * 4);
                tma_commit();
            }
        }
    }
    if (tid == 0) tma_wait_all();
    cluster.sync();                                           // nobody leaves while the peer may still touch its buffers
}

}  // namespace xfb
