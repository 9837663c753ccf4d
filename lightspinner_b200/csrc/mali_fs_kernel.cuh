// mali_fs_kernel.cuh -- the production formal-solution / Gamma kernel: fs_gamma_kernel_c<TMAX, NA>.
//
// Mapping (same as the generic fs_gamma_kernel in mali_kernels.cuh): one warp per (column, tile); lane =
// (wavelength within the tile, angle); each lane runs the downward then the upward short-characteristic recurrence
// of its ray and, at every depth point, builds the opacity / source function from the tile's transitions ("slots"),
// advances the recurrence, and contributes to J and to the Gamma integrands (warp reduce-scatter, fixed order).
//
// What makes it fast on sm_100a:
//  * warp-uniform control: a block = the SAME tile for warpsPerBlock different columns, and the tile / slot
//    descriptors travel in the kernel parameters (constant bank, indexed by blockIdx.x).  Everything that depends
//    only on the tile is therefore uniform for the compiler: branches on slot count / line-vs-continuum / first
//    touch are uniform branches, constants come from the constant bank, shared-memory offsets are immediates.
//  * nothing depth-invariant is recomputed: per slot a lane keeps one running 32-bit table index per stream and
//    adds a stride per depth step; lanes on which a transition is not active point at a zero pad with stride 0,
//    so no select is needed.
//  * the per-column depth profiles every ray needs (populations n[level][k], heights z[k]) are staged once per
//    warp into shared memory with TMA bulk copies (cp.async.bulk + mbarrier).
//  * the streaming tables (Vij, wla, background chi/eta/sca, J-dagger) of depth step s+1 are loaded into
//    registers while step s is computed (software prefetch), so ~12 resident warps per SM hide HBM latency.
//  * per-level sums of the MALI cross terms (rh_method.py:619-622) live in shared memory as one {chi, U} pair per
//    (level-slot, lane); host-computed first-touch flags turn zero-then-accumulate into store / read-modify-write.
// Arithmetic: ~29 fp64 operations per active transition and depth point + ~85 for source function, recurrence
// (one exp_m, four IEEE divides) and J, unfused and in the reference's order.
#pragma once
#include "mali_kernels.cuh"

namespace mali {

struct SlotC {  // 64 B, warp-uniform
    int32_t kind;          // bit 0: line; bit 1 / 2: first slot of the tile to touch level-slot I / J
    int32_t Nblue, Nlam;
    int32_t tabOff;        // colconst offset: lines Vij[2][N][Nlam][Nrays]; continua gij[N][Nlam]
    int32_t wlaOff;        // colconst offset: lines wla[N][Nlam]
    int32_t toff;          // offset into the per-wavelength tables (continua: alpha, twohc, wlacont)
    int32_t rowIN, rowJN;  // (row of the lower / upper level) * N in the staged populations
    int32_t lvI, lvJ;      // level-slot * 32 (double2 units) in the per-warp level array
    int32_t atom;
    int32_t pad;
    double cA, cB;         // lines: Bji/Bij, Aji/Bji
};

template <int TMAX>
struct TileC {
    int32_t la0, nslot, partRow0, pad;
    SlotC s[TMAX];
};

struct FsCommon {
    int32_t N, Nrays, Nspect, Lw;
    int32_t col0, ncol, warpsPerBlock, useBulk;
    int32_t smemBytesPerWarp, popDoubles, zOffDoubles, lvlOffDoubles, mbarOffBytes, expTabOffBytes;
    int64_t colStride, popStride, JStride, IStride, scratchStride;
    int64_t off_z, off_bbc, off_bgchi, off_bgeta, off_bgsca, off_zero;
    int64_t off_jpart, off_part;
    const double *alpha, *twohc, *wlacont, *zmu, *hw;
    const double *colconst, *pops;
    double *J, *I, *scratch;
    unsigned long long *dJbits;
    int32_t *status;
    const int32_t *done;
};

template <int TMAX>
struct ClassParams {
    static constexpr int kMaxTiles = (TMAX <= 4) ? 112 : 56;  // keeps the parameter block under 32 KB
    FsCommon c;
    TileC<TMAX> tiles[kMaxTiles];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// deterministic reduce-scatter of M = 2, 4 or 8 values per lane over the warp; the lane ends up holding the
// total of value index lane / (32 / M).  Fixed tree -> bitwise reproducible.
template <int M>
__device__ __forceinline__ double reduce_scatter(double (&v)[8], int lane)
{
    const unsigned full = 0xffffffffu;
    int off = 16;
#pragma unroll
    for (int m = M; m > 1; m >>= 1, off >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int j = 0; j < m / 2; ++j) {
            const double send = up ? v[j] : v[j + m / 2];
            const double keep = up ? v[j + m / 2] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, off);
        }
    }
#pragma unroll
    for (; off > 0; off >>= 1) v[0] = v[0] + __shfl_xor_sync(full, v[0], off);
    return v[0];
}

#ifndef MALI_MINB4
#define MALI_MINB4 3
#endif
#ifndef MALI_MINB8
#define MALI_MINB8 2
#endif

template <int TMAX, int NA>
__global__ void __launch_bounds__(128, (TMAX <= 4) ? MALI_MINB4 : MALI_MINB8) fs_gamma_kernel_c(const __grid_constant__ ClassParams<TMAX> P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FsCommon &p = P.c;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = p.col0 + blockIdx.x * p.warpsPerBlock + warp;
    if (col >= p.col0 + p.ncol) return;
    if (p.done != nullptr && p.done[col] != 0) return;

    const TileC<TMAX> &T = P.tiles[blockIdx.y];
    const int N = p.N, Nrays = p.Nrays, Nspect = p.Nspect;
    const int nslot = T.nslot;
    const int ls = lane / Nrays, mu = lane - ls * Nrays;
    const int la = T.la0 + ls;
    const bool valid = (ls < p.Lw) && (la < Nspect);
    const int laC = valid ? la : T.la0;
    const int muC = valid ? mu : 0;
    const bool leader = valid && (mu == 0);

    const double *__restrict__ cc = p.colconst + (size_t)col * p.colStride;
    double *Jcol = p.J + (size_t)col * p.JStride;
    double *scr = p.scratch + (size_t)col * p.scratchStride;
    double *Jpart = scr + p.off_jpart;
    double *part = scr + p.off_part + (size_t)T.partRow0 * N;

    // ---- stage this column's depth profiles (populations, heights) into shared memory: TMA bulk copies
    unsigned char *wbase = smem_raw + (size_t)warp * p.smemBytesPerWarp;
    double *sN = reinterpret_cast<double *>(wbase);
    double *sZ = sN + p.zOffDoubles;
    double2 *lvl = reinterpret_cast<double2 *>(sN + p.lvlOffDoubles) + lane;  // [level-slot][32] {chi, U}
    {
        const double *gN = p.pops + (size_t)col * p.popStride;
        const double *gZ = cc + p.off_z;
        if (p.useBulk) {
            const uint32_t mbar = smem_u32(wbase + p.mbarOffBytes);
            if (lane == 0) {
                const uint32_t bytesN = (uint32_t)p.popDoubles * 8u, bytesZ = (uint32_t)N * 8u;
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytesN + bytesZ)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(sN)),
                    "l"(gN), "r"(bytesN), "r"(mbar)
                    : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(sZ)),
                    "l"(gZ), "r"(bytesZ), "r"(mbar)
                    : "memory");
            }
            __syncwarp();
            uint32_t ok = 0;
            do {
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                    : "=r"(ok)
                    : "r"(mbar)
                    : "memory");
            } while (!ok);
        } else {
            for (int q = lane; q < p.popDoubles; q += 32) sN[q] = gN[q];
            for (int q = lane; q < N; q += 32) sZ[q] = gZ[q];
            __syncwarp();
        }
    }

    const double zmu = p.zmu[muC], hw = p.hw[muC];
    const double bbc0 = cc[p.off_bbc + 2 * laC], bbc1 = cc[p.off_bbc + 2 * laC + 1];
    const double fourPi = 4.0 * kPi;
    const int zeroIdx = (int)p.off_zero;

    // ---- depth-invariant per-lane slot state
    unsigned actM = 0u;
    double ca[TMAX], cb[TMAX], cw[TMAX];  // continua: alpha, 2hc/lambda^3, wlamu (per lane); lines: unused
#pragma unroll
    for (int tt = 0; tt < TMAX; ++tt) {
        ca[tt] = 0.0;
        cb[tt] = 0.0;
        cw[tt] = 0.0;
        if (tt < nslot) {
            const SlotC &s = T.s[tt];
            const int lt = laC - s.Nblue;
            const bool act = valid && lt >= 0 && lt < s.Nlam;
            if (act) actM |= 1u << tt;
            if (!(s.kind & 1) && act) {
                ca[tt] = __ldg(p.alpha + s.toff + lt);
                cb[tt] = __ldg(p.twohc + s.toff + lt);
                cw[tt] = (__ldg(p.wlacont + s.toff + lt) * hw) * fourPi;
            }
        }
    }

    for (int d = 0; d < 2; ++d) {
        const int dk = d ? -1 : 1;
        const int kS = d ? N - 1 : 0;
        // running table indices of this direction (inactive lanes: zero pad, stride 0)
        int ia[TMAX], sa[TMAX], ib[TMAX], sb[TMAX];
#pragma unroll
        for (int tt = 0; tt < TMAX; ++tt) {
            ia[tt] = zeroIdx;
            sa[tt] = 0;
            ib[tt] = zeroIdx;
            sb[tt] = 0;
            if (tt < nslot) {
                const SlotC &s = T.s[tt];
                const int lt = laC - s.Nblue;
                if ((actM >> tt) & 1u) {
                    if (s.kind & 1) {
                        const int strA = s.Nlam * Nrays;
                        ia[tt] = s.tabOff + lt * Nrays + muC + (d * N + kS) * strA;
                        sa[tt] = dk * strA;
                        ib[tt] = s.wlaOff + lt + kS * s.Nlam;
                        sb[tt] = dk * s.Nlam;
                    } else {
                        ia[tt] = s.tabOff + lt + kS * s.Nlam;
                        sa[tt] = dk * s.Nlam;
                    }
                }
            }
        }
        int kl = kS * Nspect + laC;
        const int dkl = dk * Nspect;

        // thermalised lower boundary needs chi at kS+dk before the sweep starts (formal_solver.py:205)
        double chiProbe = 0.0;
        if (d) {
            const int k = kS + dk;
            double chiTot = 0.0;
#pragma unroll
            for (int tt = 0; tt < TMAX; ++tt) {
                if (tt < nslot) {
                    const SlotC &s = T.s[tt];
                    const double ld = __ldg(cc + ia[tt] + sa[tt]);
                    const double ni = sN[s.rowIN + k], nj = sN[s.rowJN + k];
                    if (s.kind & 1) {
                        chiTot += ni * ld - nj * (s.cA * ld);
                    } else {
                        chiTot += ni * ca[tt] - nj * (ld * ca[tt]);
                    }
                }
            }
            chiProbe = chiTot + __ldg(cc + p.off_bgchi + kl + dkl);
        }

        // software prefetch: the streams of step s+1 are in flight while step s is computed
        double ldN[TMAX], wlN[TMAX], bgcN, bgeN, bgsN, JdN;
#pragma unroll
        for (int tt = 0; tt < TMAX; ++tt) {
            ldN[tt] = 0.0;
            wlN[tt] = 0.0;
            if (tt < nslot) {
                ldN[tt] = __ldg(cc + ia[tt]);
                if (T.s[tt].kind & 1) wlN[tt] = __ldg(cc + ib[tt]);
            }
        }
        bgcN = __ldg(cc + p.off_bgchi + kl);
        bgeN = __ldg(cc + p.off_bgeta + kl);
        bgsN = __ldg(cc + p.off_bgsca + kl);
        JdN = Jcol[kl];

        Sweep sw;
        for (int s = 0; s < N; ++s) {
            const int k = kS + s * dk;
            double ld[TMAX], wl[TMAX];
#pragma unroll
            for (int tt = 0; tt < TMAX; ++tt) {
                ld[tt] = ldN[tt];
                wl[tt] = wlN[tt];
            }
            const double bgc = bgcN, bge = bgeN, bgs = bgsN, Jdag = JdN;
            const int klc = kl;
            // partial sums written by the down sweep (read early: consumed at the end of the step)
            double jOld = 0.0, gOld[(TMAX + 3) / 4];
#pragma unroll
            for (int c = 0; c < (TMAX + 3) / 4; ++c) gOld[c] = 0.0;
            if (d) {
                if (leader) jOld = __ldcg(Jpart + klc);
#pragma unroll
                for (int c = 0; c < (TMAX + 3) / 4; ++c) {
                    const int tt = 4 * c + (lane >> 3);
                    if (4 * c < nslot && (lane & 3) == 0 && tt < nslot)
                        gOld[c] = __ldcg(part + (size_t)(2 * tt + ((lane >> 2) & 1)) * N + k);
                }
            }
            if (s + 1 < N) {
                kl += dkl;
#pragma unroll
                for (int tt = 0; tt < TMAX; ++tt) {
                    if (tt < nslot) {
                        ia[tt] += sa[tt];
                        ldN[tt] = __ldg(cc + ia[tt]);
                        if (T.s[tt].kind & 1) {
                            ib[tt] += sb[tt];
                            wlN[tt] = __ldg(cc + ib[tt]);
                        }
                    }
                }
                bgcN = __ldg(cc + p.off_bgchi + kl);
                bgeN = __ldg(cc + p.off_bgeta + kl);
                bgsN = __ldg(cc + p.off_bgsca + kl);
                JdN = Jcol[kl];
            }

            // ---- (1) opacity / emissivity, rh_method.py:601-632
            double chiTot = 0.0, etaTot = 0.0;
            double eta0 = 0.0, eta1 = 0.0, eta2 = 0.0, eta3 = 0.0;  // per-atom emissivity (scalars: stay in registers)
#pragma unroll
            for (int tt = 0; tt < TMAX; ++tt) {
                if (tt < nslot) {
                    const SlotC &sl = T.s[tt];
                    double Vij, Vji, Uji;
                    if (sl.kind & 1) {  // rh_method.py:278-281; the table holds hc/4pi*Bij*phi
                        Vij = ld[tt];
                        Vji = sl.cA * Vij;
                        Uji = sl.cB * Vji;
                    } else {            // rh_method.py:284-286
                        Vij = ca[tt];
                        Vji = ld[tt] * Vij;
                        Uji = cb[tt] * Vji;
                    }
                    const double ni = sN[sl.rowIN + k], nj = sN[sl.rowJN + k];
                    const double chi_t = ni * Vij - nj * Vji;
                    const double eta_t = nj * Uji;
                    double2 *pI = lvl + sl.lvI, *pJ = lvl + sl.lvJ;
                    if (sl.kind & 2) {
                        *pI = make_double2(0.0 + chi_t, 0.0);
                    } else {
                        pI->x = pI->x + chi_t;
                    }
                    if (sl.kind & 4) {
                        *pJ = make_double2(0.0 - chi_t, 0.0 + Uji);
                    } else {
                        double2 v = *pJ;
                        v.x = v.x - chi_t;
                        v.y = v.y + Uji;
                        *pJ = v;
                    }
                    if (NA > 1 && sl.atom == 1)
                        eta1 += eta_t;
                    else if (NA > 2 && sl.atom == 2)
                        eta2 += eta_t;
                    else if (NA > 3 && sl.atom == 3)
                        eta3 += eta_t;
                    else
                        eta0 += eta_t;
                    chiTot += chi_t;
                    etaTot += eta_t;
                }
            }
            chiTot += bgc;
            const double rchi = rcp_full(chiTot);
            const double S = div_by(etaTot + bge + bgs * Jdag, chiTot, rchi);

            // ---- (2) short characteristic
            const double zk = sZ[k];
            double Ik, Psi;
            if (s == 0)
                sw.first(d != 0, zmu, chiTot, S, zk, chiProbe, sZ[kS + dk], bbc0, bbc1, Ik, Psi);
            else
                sw.step(s == N - 1, zmu, chiTot, rchi, S, zk, Ik, Psi);

            // ---- (3) J, rh_method.py:640
            {
                const double x = valid ? hw * Ik : 0.0;
                double sum = x;
                if (Nrays == 5) {  // the reference's quadrature(5): fully unrolled
#pragma unroll
                    for (int m = 1; m < 5; ++m) sum += __shfl_down_sync(0xffffffffu, x, m);
                } else if (Nrays == 3) {
#pragma unroll
                    for (int m = 1; m < 3; ++m) sum += __shfl_down_sync(0xffffffffu, x, m);
                } else {
                    for (int m = 1; m < Nrays; ++m) sum += __shfl_down_sync(0xffffffffu, x, m);
                }
                // down: store the partial; up: complete it.  j_finish_kernel then forms dJ and moves Jpart -> J
                if (leader) __stcg(Jpart + klc, d == 0 ? sum : jOld + sum);
            }

            // ---- (4) Gamma integrands, rh_method.py:643-681
#pragma unroll
            for (int c = 0; c < (TMAX + 3) / 4; ++c) {
                const int c0 = 4 * c;
                if (c0 < nslot) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int tt = c0 + q;
                        v[2 * q] = 0.0;
                        v[2 * q + 1] = 0.0;
                        if (tt < TMAX && tt < nslot) {
                            const SlotC &sl = T.s[tt];
                            double Vij, Vji, Uji, wlamu;
                            if (sl.kind & 1) {
                                Vij = ld[tt];
                                Vji = sl.cA * Vij;
                                Uji = sl.cB * Vji;
                                wlamu = (wl[tt] * hw) * fourPi;  // rh_method.py:665
                            } else {
                                Vij = ca[tt];
                                Vji = ld[tt] * Vij;
                                Uji = cb[tt] * Vji;
                                wlamu = cw[tt];
                            }
                            const double2 LI = lvl[sl.lvI], LJ = lvl[sl.lvJ];
                            double eta_a = eta0;
                            if (NA > 1 && sl.atom == 1) eta_a = eta1;
                            if (NA > 2 && sl.atom == 2) eta_a = eta2;
                            if (NA > 3 && sl.atom == 3) eta_a = eta3;
                            const double Ieff = Ik - Psi * eta_a;
                            const double g1 = (Uji + Vji * Ieff) - ((LI.x * Psi) * LJ.y);
                            const double g2 = (Vij * Ieff) - ((LJ.x * Psi) * LI.y);
                            v[2 * q] = g1 * wlamu;      // inactive lanes: wlamu == 0
                            v[2 * q + 1] = g2 * wlamu;
                        }
                    }
                    const double tot = reduce_scatter<8>(v, lane);
                    const int tt = c0 + (lane >> 3);
                    if ((lane & 3) == 0 && tt < nslot) {
                        double *dst = part + (size_t)(2 * tt + ((lane >> 2) & 1)) * N + k;
                        __stcg(dst, d == 0 ? tot : gOld[c] + tot);
                    }
                }
            }
        }
        if (d == 1 && valid) p.I[(size_t)col * p.IStride + (size_t)la * Nrays + mu] = sw.Iupw;
        if (valid && sw.bad && p.status != nullptr) atomicOr(p.status + col, 2);
    }

}

}  // namespace mali
