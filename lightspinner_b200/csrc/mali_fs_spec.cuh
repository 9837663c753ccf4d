// mali_fs_spec.cuh -- structure-specialised formal-solution / Gamma kernel (fs_body<SPEC>, launched through the
// three "mega" kernels fs_gamma_kernel_m<CLS> of mali_api.cu).
//
// Same mapping and arithmetic as the generic fs_gamma_kernel (mali_kernels.cuh), but the STRUCTURE of the tile --
// how many transitions overlap it, which are lines, which atom and which lower / upper level each one connects
// (i.e. which transitions share a level in the MALI cross terms, rh_method.py:619-622, 677-680) -- is a
// compile-time constant (a C++20 constexpr struct carried by a tag type).  Consequences on sm_100a:
//   * the per-level sums chi[level], U[level] and the per-atom emissivity are plain registers with compile-time
//     indices: no shared-memory read-modify-write, no first-touch branches, no address arithmetic;
//   * slot loops unroll to the exact transition count; line / continuum code is selected at compile time;
//   * the Gamma reduce-scatter is sized to the exact number of matrix entries of the tile (2, 4, 8 or 16 values).
// What stays run-time (warp-uniform, constant bank): where the tile sits (first wavelength, table offsets, Nblue,
// Nlambda per transition, the lines' Einstein ratios) -- so one instance serves every tile with that structure.
//
// Work split: one warp per (column, tile, sweep direction) -- the down and the up sweep of a tile are independent
// until their partial sums are added by the finish kernels.
// Data movement: the depth loop contains NO global loads.  A 3-stage TMA ring (cp.async.bulk + mbarrier, issued
// by one elected lane from warp-uniform registers) streams each depth step's tile record -- this direction's Vij
// rows, the per-wavelength fields and J-dagger -- into shared memory two steps ahead; the column's heights and the
// level populations the tile touches are staged by TMA bulk copies before the sweep.  The down and the up sweep
// store their J / Gamma partial sums to separate scratch copies (no read-modify-write); the second half of each
// step's warp reductions runs during the following step so that its latency overlaps arithmetic.
//
// Instances for the structures of known models are generated ahead of time (tools/gen_spec_instances.py ->
// spec_instances.inc, compiled by nvcc into libmali_b200.so); tiles whose structure has no instance run on the
// generic kernel.
#pragma once
#include "mali_kernels.cuh"

namespace mali {

constexpr int kSpecMaxSlots = 8;
#ifndef MALI_NST
#define MALI_NST 3
#endif
constexpr int kRingStages = MALI_NST;   // stages of the per-warp TMA ring (shared with the host's smem sizing)

struct FsCommon {  // launch-invariant parameters of the specialised kernels (constant bank)
    int32_t N, Nrays, Nspect, Lw;
    int32_t col0, ncol, warpsPerBlock, useBulk;
    int32_t smemBytesPerWarp, popDoubles, zOffDoubles, lvlOffDoubles, mbarOffBytes, expTabOffBytes;
    int32_t ringOffDoubles, pad1;
    int64_t colStride, popStride, JStride, IStride, scratchStride;
    int64_t off_z, off_bbc, off_tab, rowStride;
    int64_t off_jpart, off_part, upOff;
    const double *alpha, *twohc, *wlacont, *zmu, *hw;
    const double *colconst, *pops;
    double *J, *I, *scratch;
    unsigned long long *dJbits;
    int32_t *status;
    const int32_t *done;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct TileStruct {
    int lw;                       // wavelengths per tile (32 / Nrays): fixes every offset inside a record
    int nslot;                    // transitions overlapping the tile
    int natom;                    // atoms of the model (per-atom emissivity registers)
    int nlev;                     // distinct (atom, level) pairs touched
    int kind[kSpecMaxSlots];      // 1 = line, 0 = continuum
    int atom[kSpecMaxSlots];
    int lvI[kSpecMaxSlots], lvJ[kSpecMaxSlots];    // level-slot of the lower / upper level
    int rowI[kSpecMaxSlots], rowJ[kSpecMaxSlots];  // row of the level in n[sumNlevel][N]
    int nrays;                    // angles per wavelength (the mu-sum of J unrolls exactly)
};

struct SlotR {  // run-time part of a slot, warp-uniform (constant bank), 32 B
    int32_t Nblue, Nlam, toff, pad;
    double cA, cB;  // lines: Bji/Bij, Aji/Bji
};

// Tiles are grouped into three register classes by slot count; one "mega" kernel per class dispatches on the
// tile's structure id (a warp-uniform switch), so a whole class is ONE launch however many structures it holds.
__host__ __device__ constexpr int spec_class(int nslot) { return nslot <= 2 ? 0 : (nslot <= 4 ? 1 : 2); }
__host__ __device__ constexpr int spec_class_slots(int cls) { return cls == 0 ? 2 : (cls == 1 ? 4 : 8); }
// resident warps per SM requested from ptxas (blocks are single warps): sets the register budget per class
#ifndef MALI_OCC0
#define MALI_OCC0 16
#endif
#ifndef MALI_OCC1
#define MALI_OCC1 14
#endif
#ifndef MALI_OCC2
#define MALI_OCC2 12
#endif
__host__ __device__ constexpr int spec_class_warps(int cls) { return cls == 0 ? MALI_OCC0 : (cls == 1 ? MALI_OCC1 : MALI_OCC2); }
// number of line slots among the first tt slots (tt == nslot: all of them) -> position of a line's Vij rows
__host__ __device__ constexpr int spec_line_index(const TileStruct &S, int tt)
{
    int n = 0;
    for (int u = 0; u < tt; ++u) n += S.kind[u] ? 1 : 0;
    return n;
}
__host__ __device__ constexpr int spec_pow2(int x) { return x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : 16)); }

template <int NSP>
struct TileR {
    int32_t la0, partRow0, spec, recOff;  // spec: structure id (index of the ahead-of-time instance)
    SlotR s[NSP];
};

template <int CLS>
struct MegaParams {
    static constexpr int NSP = spec_class_slots(CLS);
    static constexpr int kMaxTiles = (31 * 1024 - (int)sizeof(FsCommon)) / (int)sizeof(TileR<NSP>);
    FsCommon c;
    TileR<NSP> tiles[kMaxTiles];
};

// Warp reduce-scatter through shared memory, split in two so that the second half can run one depth step later
// (its latency then overlaps the next step's arithmetic): reduce_store puts the lane's M values into the per-warp
// scratch (M rows x 36 doubles; lane l of a row sits at l + l/8: conflict-free stores and loads); reduce_load sums
// them: afterwards the lane holds the total over the 32 lanes of value index L / (32/M).  Fixed order -> deterministic.
template <int M>
__device__ __forceinline__ void reduce_store(const double (&v)[M], int lane, double *red)
{
#pragma unroll
    for (int q = 0; q < M; ++q) red[q * 36 + lane + (lane >> 3)] = v[q];
}
template <int M>
__device__ __forceinline__ double reduce_load(int lane, const double *red)
{
    constexpr int G = 32 / M;      // lanes that share one value after the reduction
    constexpr int SEG = 32 / G;    // source lanes each of them sums
    const int e = lane / G, part = lane % G;
    const double *row = red + e * 36;
    double acc = 0.0, acc1 = 0.0;   // two interleaved partial sums halve the dependent-add chain
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        const int l = part * SEG + i;
        const double x = row[l + (l >> 3)];
        if (i & 1)
            acc1 = (i == 1) ? x : acc1 + x;
        else
            acc = (i == 0) ? x : acc + x;
    }
    if (SEG > 1) acc = acc + acc1;
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
    return acc;
}

// row of n[sumNlevel][N] that holds level-slot lv of the tile
__host__ __device__ constexpr int spec_level_row(const TileStruct &S, int lv)
{
    for (int u = 0; u < S.nslot; ++u) {
        if (S.lvI[u] == lv) return S.rowI[u];
        if (S.lvJ[u] == lv) return S.rowJ[u];
    }
    return 0;
}

// populations of the levels a tile touches -> sN[level-slot][depth] (plain loads: N odd, no bulk copy possible)
template <class SPEC, int LV>
__device__ __forceinline__ void stage_levels(double *sN, const double *gN, int N, int lane)
{
    constexpr TileStruct S = SPEC::S;
    if constexpr (LV < S.nlev) {
        constexpr int row = spec_level_row(S, LV);   // constant-evaluated: S never materialises in local memory
        const double *src = gN + (size_t)row * N;
        for (int k = lane; k < N; k += 32) sN[LV * N + k] = src[k];
        stage_levels<SPEC, LV + 1>(sN, gN, N, lane);
    }
}

#define NIDX(k, lv) ((lv) * N + (k))   // populations in shared memory: [level-slot][depth]
// one TMA bulk copy per level row the tile touches -> sN[level-slot][depth]
template <class SPEC, int LV>
__device__ __forceinline__ void stage_levels_bulk(double *sN, const double *gN, int N, uint32_t mbar)
{
    constexpr TileStruct S = SPEC::S;
    if constexpr (LV < S.nlev) {
        constexpr int row = spec_level_row(S, LV);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sN + (size_t)LV * N)),
                     "l"(gN + (size_t)row * N), "r"((uint32_t)N * 8u), "r"(mbar)
                     : "memory");
        stage_levels_bulk<SPEC, LV + 1>(sN, gN, N, mbar);
    }
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}

// TMA ring producer step: one elected lane fetches the record of sweep step fetchS (0 .. N-1 in sweep order) into
// the next ring stage -- this direction's Vij rows and the per-wavelength fields -- then the (warp-uniform)
// bookkeeping advances.  A record: [Vij rows dir 0][Vij rows dir 1][fields]; VBLK / SMALL in doubles; dirOff selects
// this direction's Vij rows, stepRec = +-rowStride moves to the next depth of the sweep.
template <int VBLK, int SMALL, int NST>
__device__ __forceinline__ void ring_fetch(int &fetchS, uint32_t &fetchOff, uint32_t &fetchBar, const double *&fetchRec,
                                           int N, int64_t stepRec, int dirOff, uint32_t ringAddr, uint32_t barAddr)
{
    constexpr int STAGE = VBLK + SMALL;
    if (fetchS < N) {
        const double *srcV = fetchRec + dirOff;
        const double *srcS = fetchRec + 2 * VBLK;
        if (elect_one()) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fetchBar), "r"((uint32_t)(STAGE * 8))
                         : "memory");
            if (VBLK > 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 ringAddr + fetchOff),
                             "l"(srcV), "r"((uint32_t)(VBLK * 8)), "r"(fetchBar)
                             : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             ringAddr + fetchOff + (uint32_t)VBLK * 8u),
                         "l"(srcS), "r"((uint32_t)(SMALL * 8)), "r"(fetchBar)
                         : "memory");
        }
        ++fetchS;
        const bool wrap = fetchOff == (uint32_t)((NST - 1) * STAGE * 8);
        fetchOff = wrap ? 0u : fetchOff + (uint32_t)(STAGE * 8);
        fetchBar = wrap ? barAddr : fetchBar + 8u;
        fetchRec += stepRec;
    }
}

// second half of a depth step's reductions (run during the next step): Gamma partial of the lane's matrix entry
// and the mu-sum of J, rh_method.py:640.  The down and the up sweep write separate partials (no read-modify-write:
// the depth loop holds no global loads); gamma_finish_kernel / j_finish_kernel add them.
template <int M, int NS, int NR>
__device__ __forceinline__ void finish_step(int lane, const double *red, bool writer, bool leader, double *gdst,
                                            double *jdst, double x)
{
    if constexpr (NS > 0) {
        const double tot = reduce_load<M>(lane, red);
        if (writer) __stcg(gdst, tot);
    }
    double sum = x;
#pragma unroll
    for (int m = 1; m < NR; ++m) sum += __shfl_down_sync(0xffffffffu, x, m);   // the reference's order over mu
    if (leader) __stcg(jdst, sum);
}

// depth-loop unroll factor per register class (<=2, <=4, <=8 transitions)
#ifndef MALI_UNROLL0
#define MALI_UNROLL0 2
#endif
#ifndef MALI_UNROLL1
#define MALI_UNROLL1 2
#endif
#ifndef MALI_UNROLL2
#define MALI_UNROLL2 1
#endif

// SPEC is a tag type with a `static constexpr TileStruct S` member (the structure travels inside a type).
template <class SPEC, int NSP>
__device__ __forceinline__ void fs_body(const FsCommon &p, const TileR<NSP> &T, const int d, unsigned char *smem_raw)
{
    constexpr TileStruct S = SPEC::S;
    constexpr int NS = S.nslot;
    constexpr int NSA = NS > 0 ? NS : 1;
    constexpr int NLV = S.nlev > 0 ? S.nlev : 1;
    constexpr int NA = S.natom > 0 ? S.natom : 1;
    constexpr int M = spec_pow2(2 * NS);
    constexpr int kUnroll = NS > 4 ? MALI_UNROLL2 : (NS > 2 ? MALI_UNROLL1 : MALI_UNROLL0);
    // one warp per block: the column (hence every base pointer) is block-uniform -> uniform-register addressing
    const int lane = threadIdx.x;
    const int col = p.col0 + blockIdx.x;
    if (col >= p.col0 + p.ncol) return;
    if (p.done != nullptr && p.done[col] != 0) return;

    const int N = p.N, Nrays = p.Nrays, Nspect = p.Nspect;
    const int ls = lane / Nrays, mu = lane - ls * Nrays;
    const int la = T.la0 + ls;
    const bool valid = (ls < p.Lw) && (la < Nspect);
    const int laC = valid ? la : T.la0;
    const int muC = valid ? mu : 0;
    const bool leader = valid && (mu == 0);

    const double *__restrict__ cc = p.colconst + (size_t)col * p.colStride;
    double *scr = p.scratch + (size_t)col * p.scratchStride;
    double *Jpart = scr + p.off_jpart;
    double *part = scr + p.off_part + (size_t)T.partRow0 * N;

    // ---- per-warp shared memory: populations | heights | Gamma reduce scratch | barriers | exp table | TMA ring
    unsigned char *wbase = smem_raw;
    double *sN = reinterpret_cast<double *>(wbase);
    double *sZ = sN + p.zOffDoubles;
    double *red = sN + p.lvlOffDoubles;
    {
        // heights: one TMA bulk copy; populations of the levels this tile touches: transposed to [depth][level-slot]
        // (one running pointer and compile-time offsets in the depth loop), overlapped with the bulk copy
        const double *gN = p.pops + (size_t)col * p.popStride;
        const double *gZ = cc + p.off_z;
        const uint32_t mbar = smem_u32(wbase + p.mbarOffBytes);
        if (p.useBulk) {
            if (lane == 0) {
                const uint32_t bytesZ = (uint32_t)N * 8u;
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar),
                             "r"(bytesZ * (uint32_t)(1 + S.nlev))
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(sZ)),
                    "l"(gZ), "r"(bytesZ), "r"(mbar)
                    : "memory");
            }
        } else {
            for (int q = lane; q < N; q += 32) sZ[q] = gZ[q];
        }
        if constexpr (NS > 0) {
            if (p.useBulk) {
                if (lane == 0) stage_levels_bulk<SPEC, 0>(sN, gN, N, mbar);
            } else {
                stage_levels<SPEC, 0>(sN, gN, N, lane);
            }
        }
        __syncwarp();
        if (p.useBulk) {
            uint32_t ok = 0;
            do {
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                    : "=r"(ok)
                    : "r"(mbar)
                    : "memory");
            } while (!ok);
        }
    }

    // the exp table (2 KB) in shared memory: one 16-byte LDS per exp instead of a global load in the critical chain
    ulonglong2 *stab = reinterpret_cast<ulonglong2 *>(wbase + p.expTabOffBytes);
#pragma unroll
    for (int q = 0; q < 4; ++q) stab[lane + 32 * q] = kExpTab[lane + 32 * q];
    __syncwarp();

    const double zmu = p.zmu[muC], hw = p.hw[muC];
    const double bbc0 = cc[p.off_bbc + 2 * laC], bbc1 = cc[p.off_bbc + 2 * laC + 1];
    const double fourPi = 4.0 * kPi;
    const double r3 = rcp_full(3.0);

    // ---- compile-time record geometry (mali_types.cuh): [Vij rows dir 0][Vij rows dir 1][bg chi|eta|sca][slot fields]
    constexpr int LW = S.lw;
    constexpr int NLINE = spec_line_index(S, S.nslot);
    constexpr int VBLK = NLINE * kVRow;                                    // one direction's Vij rows
    constexpr int JW = (LW + 3) / 4 * 4;                                   // J-dagger field (whole sectors)
    constexpr int SMALL = (JW + (3 + NS) * LW + 15) / 16 * 16;             // J-dagger + bg + slot fields, padded as packed
    constexpr int STAGE = VBLK + SMALL;                                    // doubles of one ring stage
    constexpr int NST = kRingStages;                                       // ring depth: NST - 1 steps in flight
#define line_index(tt) spec_line_index(S, (tt))
    // idle lanes (lane >= Lw * Nrays) read the zero padding of the Vij rows and the last wavelength's fields; lanes
    // past the end of the spectrum read the zero / clamped entries packed for them.  Their weights are zero.
    const int lsC = ls < p.Lw ? ls : p.Lw - 1;
    const double hwG = valid ? hw : 0.0;
    const double *tab0 = cc + p.off_tab + T.recOff;   // this tile's record in depth row 0

    // ---- TMA ring: the record of depth step g+2 streams into shared memory while step g is computed.
    // All ring bookkeeping is warp-uniform (uniform registers); one elected lane issues two bulk copies per step
    // (this direction's Vij rows; the per-wavelength fields) onto the stage's mbarrier and every lane waits on it
    // before reading.  No per-lane global loads, no prefetch registers.
    double *ring = sN + p.ringOffDoubles;
    const uint32_t ringAddr = smem_u32(ring);
    const uint32_t barAddr = smem_u32(wbase + p.mbarOffBytes) + 8;  // the ring's barriers follow the staging barrier
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NST; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr + 8 * q) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // this warp's sweep: d == 0 downwards from the top (k = 0), d == 1 upwards from the bottom (k = N - 1)
    const int dk = d ? -1 : 1;
    const int kS = d ? N - 1 : 0;
    const int64_t stepRec = d ? -p.rowStride : p.rowStride;
    int fetchS = 0;                     // next sweep step to fetch
    uint32_t fetchOff = 0;              // byte offset of its stage in the ring
    uint32_t fetchBar = barAddr;        // its barrier
    const double *fetchRec = tab0 + (size_t)kS * p.rowStride;      // its record
#define fetch_next() ring_fetch<VBLK, SMALL, NST>(fetchS, fetchOff, fetchBar, fetchRec, N, stepRec, d * VBLK, ringAddr, barAddr)
#pragma unroll
    for (int q = 0; q < NST - 1; ++q) fetch_next();
    uint32_t useOff = 0, useBar = barAddr, phases = 0, useBit = 1;   // stage being consumed; parity bit per stage

    // ---- depth-invariant per-lane constants of the continuum slots (alpha, 2hc/lambda^3, wlamu)
    double ca[NSA], cb[NSA], cw[NSA];
#pragma unroll
    for (int tt = 0; tt < NS; ++tt) {
        const SlotR &s = T.s[tt];
        const int lt = laC - s.Nblue;
        const bool act = valid && lt >= 0 && lt < s.Nlam;
        ca[tt] = 0.0;
        cb[tt] = 0.0;
        cw[tt] = 0.0;
        if (!S.kind[tt] && act) {
            ca[tt] = __ldg(p.alpha + s.toff + lt);
            cb[tt] = __ldg(p.twohc + s.toff + lt);
            cw[tt] = (__ldg(p.wlacont + s.toff + lt) * hw) * fourPi;
        }
    }

    // Gamma value owned by this lane after the reduce, and where its partial sums live
    const int eOwn = lane / (32 / M);
    const bool writer = (lane % (32 / M)) == 0 && eOwn < 2 * NS;
    double *gbase = part + (size_t)eOwn * N;

    {
        int kl = kS * Nspect + laC;
        const int dkl = dk * Nspect;

        // thermalised lower boundary needs chi at kS+dk before the sweep starts (formal_solver.py:205)
        double chiProbe = 0.0;
        if (d) {
            const int k = kS + dk;
            const double *rec = tab0 + (size_t)k * p.rowStride;
            double chiTot = 0.0;
#pragma unroll
            for (int tt = 0; tt < NS; ++tt) {
                const double ni = sN[NIDX(k, S.lvI[tt])], nj = sN[NIDX(k, S.lvJ[tt])];
                if (S.kind[tt]) {
                    const double ld = __ldg(rec + VBLK + line_index(tt) * kVRow + lane);
                    chiTot += ni * ld - nj * (T.s[tt].cA * ld);
                } else {
                    const double ld = __ldg(rec + 2 * VBLK + JW + (3 + tt) * LW + lsC);
                    chiTot += ni * ca[tt] - nj * (ld * ca[tt]);
                }
            }
            chiProbe = chiTot + __ldg(rec + 2 * VBLK + JW + lsC);
        }

        SweepT<1> sw;
        sw.r3 = r3;
        sw.stab = stab;

        // the second half of a step's J / Gamma reductions runs during the NEXT step (its shuffle / shared-memory
        // latency then overlaps that step's arithmetic); state carried for it:
        double xP = 0.0;
        int kP = kS, klP = kl;
        double *gdir = gbase + (d ? p.upOff : 0), *jdir = Jpart + (d ? p.upOff : 0);
#define finish_prev() finish_step<M, NS, S.nrays>(lane, red, writer, leader, gdir + kP, jdir + klP, xP)

        const double *nk = sN + kS;   // populations at the current depth
        {   // first point
            const int s = 0;
#define STEP_KIND 0
#include "mali_fs_step.inc"
#undef STEP_KIND
        }
#pragma unroll kUnroll
        for (int s = 1; s < N - 1; ++s) {
#define STEP_KIND 1
#include "mali_fs_step.inc"
#undef STEP_KIND
        }
        {   // last point (N >= 3 is checked at model creation)
            const int s = N - 1;
#define STEP_KIND 2
#include "mali_fs_step.inc"
#undef STEP_KIND
        }
        finish_prev();
        __syncwarp();
        if (d == 1 && valid) p.I[(size_t)col * p.IStride + (size_t)la * Nrays + mu] = sw.Iupw;
        if (valid && sw.bad && p.status != nullptr) atomicOr(p.status + col, 2);
    }
#undef line_index
#undef fetch_next
#undef NIDX
#undef finish_prev
}

// ---- registry of ahead-of-time instances -----------------------------------------------------------------
struct SpecEntry {
    const char *key;
    int nslot;
    int id;
};

}  // namespace mali
