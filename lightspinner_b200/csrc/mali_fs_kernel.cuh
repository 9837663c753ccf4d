// mali_fs_kernel.cuh -- the production formal-solution / Gamma kernel: fs_gamma_kernel_t<TMAX, NA>.
//
// Same mapping and arithmetic as the generic fs_gamma_kernel (mali_kernels.cuh) -- one warp per (column, tile),
// lane = (wavelength, angle), both depth recurrences per lane -- but everything that does not depend on depth is
// hoisted into registers: the transitions of the tile ("slots") are unrolled at compile time (TMAX of them),
// their table indices are running 32-bit offsets that advance by a stride per depth step, and their constants
// live in registers.  Tiles are grouped into classes by slot count so that a tile with 3 transitions does not pay
// for 8; tiles with more than 8 go to the generic kernel.
//
// Per depth step and lane: ~29 fp64 operations per active transition + ~85 for the source function, the short
// characteristic (one exp_m, four IEEE divides) and J -- all unfused and in the reference's order.  The per-level
// sums the MALI cross terms need (rh_method.py:619-622: atom.chi[level], atom.U[level]) are kept in shared memory
// as one {chi, U} pair per (level-slot, lane); "first touch" flags computed on the host turn the reference's
// zero-then-accumulate into store / read-modify-write so nothing has to be cleared per step.
#pragma once
#include "mali_kernels.cuh"

namespace mali {

template <int TMAX, int NA>
__global__ void __launch_bounds__(128) fs_gamma_kernel_t(const FsParams p)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tslot = (blockIdx.x % p.blocksPerCol) * p.warpsPerBlock + warp;
    const int col = p.col0 + blockIdx.x / p.blocksPerCol;
    if (tslot >= p.nClassTiles) return;
    if (p.done != nullptr && p.done[col] != 0) return;

    const TileDesc td = p.tiles[p.classTiles[tslot]];
    const int N = p.N, Nrays = p.Nrays, Nspect = p.Nspect;
    const int ls = lane / Nrays, mu = lane - ls * Nrays;
    const int la = td.la0 + ls;
    const bool valid = (ls < p.Lw) && (la < Nspect);
    const int laC = valid ? la : td.la0;
    const int muC = valid ? mu : 0;
    const bool leader = valid && (mu == 0);
    const int nslot = td.nslot;

    double2 *lvl = reinterpret_cast<double2 *>(smem + (size_t)warp * p.smemPerWarp) + lane;  // [level-slot][32] {chi, U}

    const double *__restrict__ cc = p.colconst + (size_t)col * p.colStride;
    const double *__restrict__ npop = p.pops + (size_t)col * p.popStride;
    double *Jcol = p.J + (size_t)col * p.JStride;
    double *scr = p.scratch + (size_t)col * p.scratchStride;
    double *Jpart = scr + p.off_jpart;
    double *part = scr + p.off_part + (size_t)td.partRow0 * N;

    const double zmu = p.zmu[muC], hw = p.hw[muC];
    const double bbc0 = cc[p.off_bbc + 2 * laC], bbc1 = cc[p.off_bbc + 2 * laC + 1];
    const double fourPi = 4.0 * kPi;

    // ---- depth-invariant per-slot state
    int idxA[TMAX], idxB[TMAX], strA[TMAX], strB[TMAX], rows[TMAX], meta[TMAX];
    double cA[TMAX], cB[TMAX], cC[TMAX];
    unsigned actM = 0u, lineM = 0u;
#pragma unroll
    for (int tt = 0; tt < TMAX; ++tt) {
        idxA[tt] = 0;
        idxB[tt] = 0;
        strA[tt] = 0;
        strB[tt] = 0;
        rows[tt] = 0;
        meta[tt] = 0;
        cA[tt] = 0.0;
        cB[tt] = 0.0;
        cC[tt] = 0.0;
        if (tt < nslot) {
            const SlotDesc &sd = p.slots[td.slot0 + tt];
            const int lt = laC - sd.Nblue;
            const bool act = valid && lt >= 0 && lt < sd.Nlam;
            const int ltC = act ? lt : 0;
            if (act) actM |= 1u << tt;
            if (sd.isLine) {
                lineM |= 1u << tt;
                idxA[tt] = (int)sd.tabOff + ltC * Nrays + muC;  // Vij[d][k][lt][mu]
                strA[tt] = sd.Nlam * Nrays;
                idxB[tt] = (int)sd.wlaOff + ltC;                // wla[k][lt]
                strB[tt] = sd.Nlam;
                cA[tt] = sd.c2;                                 // Vji = (Bji/Bij) * Vij
                cB[tt] = sd.c1;                                 // Uji = (Aji/Bji) * Vji
            } else {
                idxA[tt] = (int)sd.tabOff + ltC;                // gij[k][lt]
                strA[tt] = sd.Nlam;
                cA[tt] = act ? __ldg(p.alpha + sd.toff + ltC) : 0.0;              // Vij = alpha
                cB[tt] = __ldg(p.twohc + sd.toff + ltC);                           // Uji = 2hc/lambda^3 * Vji
                cC[tt] = act ? (__ldg(p.wlacont + sd.toff + ltC) * hw) * fourPi : 0.0;  // wlamu, depth-invariant
            }
            rows[tt] = sd.rowI | (sd.rowJ << 16);
            meta[tt] = sd.lsI | (sd.lsJ << 8) | (sd.flags << 16) | (sd.atom << 20);
        }
    }

    unsigned long long dJb = 0ull;

    for (int d = 0; d < 2; ++d) {
        const int dk = d ? -1 : 1;
        const int kS = d ? N - 1 : 0;
        int ia[TMAX], ib[TMAX];
#pragma unroll
        for (int tt = 0; tt < TMAX; ++tt) {
            ia[tt] = idxA[tt] + (((lineM >> tt) & 1u) ? (d * N + kS) : kS) * strA[tt];
            ib[tt] = idxB[tt] + kS * strB[tt];
        }
        // thermalised lower boundary needs chi at kS+dk before the sweep starts (formal_solver.py:205)
        double chiProbe = 0.0;
        if (d) {
            const int k = kS + dk;
            double chiTot = 0.0;
#pragma unroll
            for (int tt = 0; tt < TMAX; ++tt) {
                if (tt < nslot) {
                    const bool act = (actM >> tt) & 1u;
                    const bool isLine = (lineM >> tt) & 1u;
                    const double ld = act ? __ldg(cc + ia[tt] + dk * strA[tt]) : 0.0;
                    const double Vij = isLine ? ld : cA[tt];
                    const double Vji = cA[tt] * ld;
                    const double ni = npop[(rows[tt] & 0xffff) * N + k], nj = npop[(rows[tt] >> 16) * N + k];
                    chiTot += ni * Vij - nj * Vji;
                }
            }
            chiProbe = chiTot + __ldg(cc + p.off_bgchi + (size_t)k * Nspect + laC);
        }

        Sweep sw;
        int kl = kS * Nspect + laC;
        for (int s = 0; s < N; ++s, kl += dk * Nspect) {
            const int k = kS + s * dk;
            // ---- (1) opacity / emissivity, rh_method.py:601-632
            double chiTot = 0.0, etaTot = 0.0;
            double etaA[NA], ld[TMAX];
#pragma unroll
            for (int a = 0; a < NA; ++a) etaA[a] = 0.0;
#pragma unroll
            for (int tt = 0; tt < TMAX; ++tt) {
                ld[tt] = 0.0;
                if (tt < nslot) {
                    const bool act = (actM >> tt) & 1u;
                    const bool isLine = (lineM >> tt) & 1u;
                    ld[tt] = act ? __ldg(cc + ia[tt]) : 0.0;
                    const double Vij = isLine ? ld[tt] : cA[tt];
                    const double Vji = cA[tt] * ld[tt];
                    const double Uji = cB[tt] * Vji;
                    const double ni = npop[(rows[tt] & 0xffff) * N + k], nj = npop[(rows[tt] >> 16) * N + k];
                    const double chi_t = ni * Vij - nj * Vji;
                    const double eta_t = nj * Uji;
                    const int m = meta[tt];
                    double2 *pI = lvl + (m & 0xff) * 32, *pJ = lvl + ((m >> 8) & 0xff) * 32;
                    if (m & (1 << 16)) {
                        *pI = make_double2(0.0 + chi_t, 0.0);
                    } else {
                        pI->x = pI->x + chi_t;
                    }
                    if (m & (1 << 17)) {
                        *pJ = make_double2(0.0 - chi_t, 0.0 + Uji);
                    } else {
                        double2 v = *pJ;
                        v.x = v.x - chi_t;
                        v.y = v.y + Uji;
                        *pJ = v;
                    }
                    const int atom = m >> 20;
#pragma unroll
                    for (int a = 0; a < NA; ++a) etaA[a] += (atom == a) ? eta_t : 0.0;
                    chiTot += chi_t;
                    etaTot += eta_t;
                }
            }
            chiTot += __ldg(cc + p.off_bgchi + kl);
            const double Jdag = Jcol[kl];
            const double S = (etaTot + __ldg(cc + p.off_bgeta + kl) + __ldg(cc + p.off_bgsca + kl) * Jdag) / chiTot;

            // ---- (2) short characteristic
            const double zk = __ldg(cc + p.off_z + k);
            double Ik, Psi;
            if (s == 0)
                sw.first(d != 0, zmu, chiTot, S, zk, chiProbe, __ldg(cc + p.off_z + kS + dk), bbc0, bbc1, Ik, Psi);
            else
                sw.step(s == N - 1, zmu, chiTot, S, zk, Ik, Psi);

            // ---- (3) J, rh_method.py:640
            {
                const double x = valid ? hw * Ik : 0.0;
                double sum = x;
                for (int m = 1; m < Nrays; ++m) sum += __shfl_down_sync(0xffffffffu, x, m);
                if (leader) {
                    if (d == 0) {
                        __stcg(Jpart + kl, sum);
                    } else {
                        const double Jn = __ldcg(Jpart + kl) + sum;
                        Jcol[kl] = Jn;
                        const unsigned long long b = absbits(1.0 - Jdag / Jn);
                        dJb = b > dJb ? b : dJb;
                    }
                }
            }

            // ---- (4) Gamma integrands, rh_method.py:643-681
#pragma unroll
            for (int c0 = 0; c0 < TMAX; c0 += 4) {
                if (c0 < nslot) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int tt = c0 + q;
                        v[2 * q] = 0.0;
                        v[2 * q + 1] = 0.0;
                        if (tt < TMAX && tt < nslot) {
                            const bool act = (actM >> tt) & 1u;
                            const bool isLine = (lineM >> tt) & 1u;
                            const double Vij = isLine ? ld[tt] : cA[tt];
                            const double Vji = cA[tt] * ld[tt];
                            const double Uji = cB[tt] * Vji;
                            double wlamu = cC[tt];
                            if (isLine) wlamu = act ? (__ldg(cc + ib[tt]) * hw) * fourPi : 0.0;
                            const int m = meta[tt];
                            const double2 LI = lvl[(m & 0xff) * 32], LJ = lvl[((m >> 8) & 0xff) * 32];
                            const int atom = m >> 20;
                            double eta_a = etaA[0];
#pragma unroll
                            for (int a = 1; a < NA; ++a) eta_a = (atom == a) ? etaA[a] : eta_a;
                            const double Ieff = Ik - Psi * eta_a;
                            const double g1 = (Uji + Vji * Ieff) - ((LI.x * Psi) * LJ.y);
                            const double g2 = (Vij * Ieff) - ((LJ.x * Psi) * LI.y);
                            v[2 * q] = act ? g1 * wlamu : 0.0;
                            v[2 * q + 1] = act ? g2 * wlamu : 0.0;
                        }
                    }
                    const double tot = reduce_scatter8(v, lane);
                    const int e = lane >> 2;
                    const int tt = c0 + (e >> 1);
                    if ((lane & 3) == 0 && tt < nslot) {
                        double *dst = part + (size_t)(2 * tt + (e & 1)) * N + k;
                        if (d == 0)
                            __stcg(dst, tot);
                        else
                            __stcg(dst, __ldcg(dst) + tot);
                    }
                }
            }
#pragma unroll
            for (int tt = 0; tt < TMAX; ++tt) {
                ia[tt] += dk * strA[tt];
                ib[tt] += dk * strB[tt];
            }
        }
        if (d == 1 && valid) p.I[(size_t)col * p.IStride + (size_t)la * Nrays + mu] = sw.Iupw;
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, dJb, off);
        dJb = o > dJb ? o : dJb;
    }
    if (lane == 0) atomicMax(p.dJbits + col, dJb);
}

}  // namespace mali
