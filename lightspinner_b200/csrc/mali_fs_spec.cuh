// mali_fs_spec.cuh -- structure-specialised formal-solution / Gamma kernel (fs_body<SPEC, DIR>, launched through the
// three "mega" kernels fs_gamma_kernel_m<CLS> of mali_fs_class.cu).
//
// Same mapping and arithmetic as the generic fs_gamma_kernel (mali_kernels.cuh), but the STRUCTURE of the tile --
// how many transitions overlap it, which are lines, which atom and which lower / upper level each one connects
// (i.e. which transitions share a level in the MALI cross terms, rh_method.py:619-622, 677-680) -- and the sweep
// DIRECTION are compile-time constants.  Consequences on sm_100a:
//   * the per-level sums chi[level], U[level] and the per-atom emissivity are plain registers with compile-time
//     indices: no shared-memory read-modify-write, no first-touch branches, no address arithmetic;
//   * slot loops unroll to the exact transition count; line / continuum code is selected at compile time;
//   * the Gamma reduce-scatter is sized to the exact number of matrix entries of the tile (2 per transition);
//   * every shared-memory operand of the depth loop sits at a compile-time offset from one of three running bases.
// What stays run-time (warp-uniform, constant bank): where the tile sits (first wavelength, table offset, Nblue,
// Nlambda per transition, the lines' Einstein ratios) -- so one instance serves every tile with that structure.
//
// Work split: one warp per (column, tile, sweep direction) -- the down and the up sweep of a tile are independent
// until their partial sums are added by the finish kernels.
// Data movement: the depth loop contains NO global loads.  A TMA ring (cp.async.bulk + mbarrier, issued by one
// elected lane from warp-uniform registers) streams, two depth steps per stage, everything a step reads -- this
// direction's Vij rows and the per-wavelength fields (one contiguous piece of the tile's record), and the heights
// and level populations of those depths (two rows of the depth-major popsT table) -- so the shared memory of a warp
// does not depend on the number of depth points.  The down and the up sweep store their J / Gamma partial sums to
// separate scratch copies (no read-modify-write); the second half of each step's warp reductions runs during the
// following step so that its latency overlaps arithmetic.
//
// Instances for the structures of known models are generated ahead of time (tools/gen_spec_instances.py ->
// spec_instances.inc, compiled by nvcc into libmali_b200.so); tiles whose structure has no instance run on the
// generic kernel.
#pragma once
#include "mali_device.cuh"

namespace mali {

constexpr int kSpecMaxSlots = 8;
#ifndef MALI_NST
#define MALI_NST 3
#endif
constexpr int kRingStages = MALI_NST;   // stages of the per-warp TMA ring; a stage holds two depth steps

struct FsCommon {  // launch-invariant parameters of the specialised kernels (constant bank)
    int32_t N, Nrays, Nspect, Lw;
    int32_t col0, ncol, popsW, colChunk;   // colChunk: columns per grid slab (blockIdx.z), see mali_fs_class.cu
    int32_t dirInterleave, pad0;
    int64_t colStride, IStride, scratchStride;
    int64_t off_bbc, off_tab, off_popsT;
    int64_t off_jpart, off_part, upOff;
    const double *alpha, *twohc, *wlacont, *zmu, *hw;
    const double *colconst;
    double *I, *scratch;
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
    int rowI[kSpecMaxSlots], rowJ[kSpecMaxSlots];  // row of the level in n[sumNlevel][N] (= column 1 + row of popsT)
    int nrays;                    // angles per wavelength (the mu-sum of J unrolls exactly)
    int pw;                       // width of a popsT row: 1 + sumNlevel rounded up to 4
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

// ---- continuum groups: the bound-free transitions of a tile that share their upper level (all continua of one
// ionisation stage end on the next stage's ground level).  For such a group the emissivity is n_j * sum_t Uji_t and
// the upper level's U is sum_t Uji_t -- a sum that does not depend on the populations.  It is formed once per
// (wavelength, depth) at upload (cont_group_kernel) and travels as one more field of the record, so that the
// contracted-arithmetic kernels spend 3 instead of 12 operations per bound-free transition on the opacity stage.
// A group is identified by its first slot; groups are numbered in slot order.
__host__ __device__ constexpr int spec_group_first(const TileStruct &S, int tt)
{
    for (int u = 0; u < tt; ++u)
        if (!S.kind[u] && S.lvJ[u] == S.lvJ[tt]) return u;
    return tt;
}
__host__ __device__ constexpr int spec_group_last(const TileStruct &S, int tt)
{
    int r = tt;
    for (int u = tt + 1; u < S.nslot; ++u)
        if (!S.kind[u] && S.lvJ[u] == S.lvJ[tt]) r = u;
    return r;
}
__host__ __device__ constexpr int spec_group_index(const TileStruct &S, int tt)
{
    const int f = spec_group_first(S, tt);
    int n = 0;
    for (int u = 0; u < f; ++u)
        if (!S.kind[u] && spec_group_first(S, u) == u) ++n;
    return n;
}
__host__ __device__ constexpr int spec_ngroup(const TileStruct &S)
{
    int n = 0;
    for (int u = 0; u < S.nslot; ++u)
        if (!S.kind[u] && spec_group_first(S, u) == u) ++n;
    return n;
}
// The contracted opacity stage, planned at compile time (evaluated by the front end: one constexpr object per
// instance, read with unrolled indices inside the depth loop).  Events, in slot order: a line touches chi of both its
// levels, U of its upper level and its atom's emissivity; a continuum touches chi of its lower level; the LAST
// continuum of a group then adds the group's sums to chi / U of the upper level and to the atom's emissivity.  The
// flags say which of these writes is the first one of its accumulator (a plain store instead of an add).
struct FastPlan {
    bool firstI[kSpecMaxSlots], firstJ[kSpecMaxSlots], firstU[kSpecMaxSlots], firstA[kSpecMaxSlots];   // slot events
    bool gOpen[kSpecMaxSlots], gClose[kSpecMaxSlots];     // continuum: first / last slot of its group
    bool gFirstJ[kSpecMaxSlots], gFirstU[kSpecMaxSlots], gFirstA[kSpecMaxSlots];                      // group events
    int grp[kSpecMaxSlots];                               // continuum: index of its group
};
__host__ __device__ constexpr FastPlan make_fast_plan(const TileStruct &S)
{
    FastPlan P{};
    bool chiT[2 * kSpecMaxSlots] = {}, UT[2 * kSpecMaxSlots] = {}, atomT[8] = {};
    for (int tt = 0; tt < S.nslot; ++tt) {
        const int li = S.lvI[tt], lj = S.lvJ[tt], a = S.atom[tt];
        P.firstI[tt] = !chiT[li];
        chiT[li] = true;
        if (S.kind[tt]) {
            P.firstJ[tt] = !chiT[lj];
            P.firstU[tt] = !UT[lj];
            P.firstA[tt] = !atomT[a];
            chiT[lj] = UT[lj] = atomT[a] = true;
        } else {
            P.grp[tt] = spec_group_index(S, tt);
            P.gOpen[tt] = spec_group_first(S, tt) == tt;
            P.gClose[tt] = spec_group_last(S, tt) == tt;
            if (P.gClose[tt]) {
                P.gFirstJ[tt] = !chiT[lj];
                P.gFirstU[tt] = !UT[lj];
                P.gFirstA[tt] = !atomT[a];
                chiT[lj] = UT[lj] = atomT[a] = true;
            }
        }
    }
    return P;
}

// ---- geometry of a record / a ring stage / the per-warp shared memory, shared by the kernel and the host's sizing
// (ns: transitions of the tile, ng: its continuum groups)
__host__ __device__ constexpr int rec_jw(int lw) { return (lw + 3) / 4 * 4; }                          // J-dagger field
__host__ __device__ constexpr int rec_fields(int lw, int ns, int ng) { return (rec_jw(lw) + (3 + ns + ng) * lw + 3) / 4 * 4; }
// doubles of one ring stage: two depth steps x (one direction's Vij rows + fields), two popsT rows
__host__ __device__ constexpr int ring_stage_doubles(int nline, int lw, int ns, int ng, int pw)
{
    return 2 * (nline * kVRow + rec_fields(lw, ns, ng)) + 2 * pw;
}
// lanes that end up sharing one Gamma value after the reduce-scatter: the largest power of two G with M * G <= 32
__host__ __device__ constexpr int red_group(int m) { return m <= 2 ? 16 : (m <= 4 ? 8 : (m <= 8 ? 4 : 2)); }
constexpr int kRedRow = 36;   // doubles per row of the reduce scratch (see reduce_store)
// bytes of shared memory of one warp: ring | reduce scratch | ring barriers | exp table
__host__ __device__ constexpr int fs_smem_bytes(int nline, int lw, int ns, int ng, int pw)
{
    return (kRingStages * ring_stage_doubles(nline, lw, ns, ng, pw) + (ns > 0 ? 2 * ns * kRedRow : 0)) * 8 +
           (kRingStages * 8 + 15) / 16 * 16 + 128 * 16;
}

template <int NSP>
struct TileR {
    int32_t la0, partRow0, spec, recOff;  // spec: structure id (index of the ahead-of-time instance); recOff: depth-0 record
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
// scratch -- M rows of kRedRow = 36 doubles, lane l of a row at l + 2 (l >> 4) -- and reduce_load sums them: afterwards
// the G = red_group(M) lanes e * G .. e * G + G - 1 hold the total over the 32 lanes of value e.  Fixed order ->
// deterministic.  Bank-conflict free in both phases: a store instruction writes 16 consecutive doubles per half-warp;
// the loads are 128-bit (a lane's segment starts on a 16-byte boundary: the pad after lane 15 is TWO doubles), served
// a quarter-warp at a time, and the 8 lanes of a quarter -- rows e = stride 18 sixteen-byte banks = 2 mod 8, segment
// starts {0, 4, 9, 13}, {0, 9} ... -- fall into 8 distinct banks.
template <int M>
__device__ __forceinline__ void reduce_store(const double (&v)[M], int lane, double *red)
{
    double *dst = red + lane + 2 * (lane >> 4);
#pragma unroll
    for (int q = 0; q < M; ++q) dst[q * kRedRow] = v[q];
}
template <int M>
__device__ __forceinline__ double reduce_load(int lane, const double *red)
{
    constexpr int G = red_group(M);   // lanes that share one value after the reduction
    constexpr int SEG = 32 / G;       // source lanes each of them sums (2, 4, 8 or 16)
    int e = lane / G;
    const int part = lane % G;
    if (e > M - 1) e = M - 1;         // lanes beyond M * G idle (their result is not used)
    const int l0 = part * SEG;
    const double2 *row = reinterpret_cast<const double2 *>(red + e * kRedRow + l0 + 2 * (l0 >> 4));   // SEG <= 16: no straddling
    double acc = 0.0, acc1 = 0.0;   // two interleaved partial sums halve the dependent-add chain
#pragma unroll
    for (int i = 0; i < SEG / 2; ++i) {
        const double2 x = row[i];
        acc = (i == 0) ? x.x : acc + x.x;
        acc1 = (i == 0) ? x.y : acc1 + x.y;
    }
    acc = acc + acc1;
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
    return acc;
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load(uint32_t dst, const double *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}

// SPEC is a tag type with a `static constexpr TileStruct S` member (the structure travels inside a type).
// DIR: 0 = downward sweep from the top (k = 0), 1 = upward sweep from the bottom (k = N - 1).
// FAST: arithmetic mode (mali_device.cuh, Arith): false = the reference's rounding, true = contracted.
template <class SPEC, int NSP, int DIR, bool FAST>
__device__ __forceinline__ void fs_body(const FsCommon &p, const TileR<NSP> &T, unsigned char *smem_raw)
{
    using A = Arith<FAST>;
    constexpr TileStruct S = SPEC::S;
    constexpr int NS = S.nslot;
    constexpr int NSA = NS > 0 ? NS : 1;
    constexpr int NLV = S.nlev > 0 ? S.nlev : 1;
    constexpr int NA = S.natom > 0 ? S.natom : 1;
    constexpr int NR = S.nrays;
    constexpr int M = NS > 0 ? 2 * NS : 1;          // Gamma values of the tile: [i,j] and [j,i] of every transition
    constexpr int RG = red_group(M);
    // ---- compile-time geometry (mali_types.cuh): record = [Vij rows dir 0 | fields | Vij rows dir 1]
    constexpr int LW = S.lw;
    constexpr int NLINE = spec_line_index(S, S.nslot);
    constexpr int VB = NLINE * kVRow;                // one direction's Vij rows
    constexpr int JW = rec_jw(LW);                   // J-dagger field (whole sectors)
    constexpr int NGR = spec_ngroup(S);              // continuum groups (their U sums follow the slot fields)
    constexpr int NGA = NGR > 0 ? NGR : 1;
    constexpr FastPlan PL = make_fast_plan(S);
    constexpr int SF = rec_fields(LW, NS, NGR);      // all per-wavelength fields
    constexpr int REC = 2 * VB + SF;                 // record stride (one depth point)
    constexpr int ST1 = VB + SF;                     // what one depth step of this direction reads: one contiguous piece
    constexpr int VOFF = DIR ? SF : 0;               // ... inside which the Vij rows sit here
    constexpr int FOFF = DIR ? 0 : VB;               // ... and the fields here
    constexpr int PW = S.pw;
    constexpr int STAGE = 2 * ST1 + 2 * PW;          // doubles of a ring stage (two depth steps)
    constexpr int NST = kRingStages;
    constexpr int dk = DIR ? -1 : 1;
#define line_index(tt) spec_line_index(S, (tt))

    // one warp per block: the column (hence every base pointer) is block-uniform -> uniform-register addressing
    const int lane = threadIdx.x;
    const int col = p.col0 + blockIdx.z * p.colChunk + blockIdx.x;
    if (col >= p.col0 + p.ncol) return;
    if (p.done != nullptr && p.done[col] != 0) return;

    const int N = p.N, Nspect = p.Nspect;
    const int ls = lane / NR, mu = lane - ls * NR;
    const int la = T.la0 + ls;
    const bool valid = (ls < LW) && (la < Nspect);
    const int laC = valid ? la : T.la0;
    const int muC = valid ? mu : 0;
    const bool leader = valid && (mu == 0);
    // idle lanes (lane >= Lw * Nrays) read the zero padding of the Vij rows and the last wavelength's fields; lanes
    // past the end of the spectrum read the zero / clamped entries packed for them.  Their weights are zero.
    const int lsC = ls < LW ? ls : LW - 1;

    const double *__restrict__ cc = p.colconst + (size_t)col * p.colStride;
    double *scr = p.scratch + (size_t)col * p.scratchStride + (DIR ? p.upOff : 0);
    double *jp = scr + p.off_jpart + laC + (size_t)(DIR ? p.N - 1 : 0) * p.Nspect;   // J partial of the current depth
    // ---- per-warp shared memory: TMA ring | Gamma reduce scratch | ring barriers | exp table
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *red = ring + NST * STAGE;
    constexpr int kBarOff = (NST * STAGE + (NS > 0 ? M * kRedRow : 0)) * 8;
    const uint32_t barAddr = smem_u32(smem_raw + kBarOff);
    ulonglong2 *stab = reinterpret_cast<ulonglong2 *>(smem_raw + kBarOff + (NST * 8 + 15) / 16 * 16);
    const uint32_t ringAddr = smem_u32(ring);

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NST; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr + 8 * q) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the exp table (2 KB) in shared memory: one 16-byte LDS per exp instead of a global load in the critical chain
#pragma unroll
    for (int q = 0; q < 4; ++q) stab[lane + 32 * q] = kExpTab[lane + 32 * q];
    __syncwarp();

    // ---- TMA producer state (warp-uniform).  Group g = sweep steps 2g, 2g + 1 = depths k0 = kS + 2 g dk, k0 + dk.
    // A stage receives [piece(k0) | piece(k0 + dk) | popsT rows of the two depths in ascending depth order].
    const int NG = (N + 1) >> 1;                     // groups; the last one holds a single step when N is odd
    const double *recBase = cc + p.off_tab + T.recOff + (DIR ? VB : 0);
    const double *fRec = recBase + (size_t)(DIR ? N - 1 : 0) * REC;           // piece of the next group's first depth
    const double *fPop = cc + p.off_popsT + (size_t)(DIR ? N - 2 : 0) * PW;   // lower of the next group's two popsT rows
    int fG = 0;                                      // next group to fetch
    uint32_t fOff = 0, fBar = barAddr;               // its stage (byte offset in the ring) and barrier
    auto fetch_group = [&]() {
        if (fG < NG) {
            const bool full = 2 * fG + 1 < N;
            if (elect_one()) {
                const uint32_t dst = ringAddr + fOff;
                if (full) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fBar), "r"((uint32_t)(STAGE * 8))
                                 : "memory");
                    tma_load(dst, fRec, ST1 * 8, fBar);
                    tma_load(dst + ST1 * 8, fRec + dk * REC, ST1 * 8, fBar);
                    tma_load(dst + 2 * ST1 * 8, fPop, 2 * PW * 8, fBar);
                } else {   // odd N: the last group is one depth point (k = N - 1 going down, k = 0 going up)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fBar),
                                 "r"((uint32_t)((ST1 + PW) * 8))
                                 : "memory");
                    tma_load(dst, fRec, ST1 * 8, fBar);
                    tma_load(dst + (2 * ST1 + (DIR ? PW : 0)) * 8, fPop + (DIR ? PW : 0), PW * 8, fBar);
                }
            }
            ++fG;
            const bool wrap = fOff == (uint32_t)((NST - 1) * STAGE * 8);
            fOff = wrap ? 0u : fOff + (uint32_t)(STAGE * 8);
            fBar = wrap ? barAddr : fBar + 8u;
            fRec += 2 * dk * REC;
            fPop += 2 * dk * PW;
        }
    };
#pragma unroll
    for (int q = 0; q < NST - 1; ++q) fetch_group();

    // ---- consumer state: stage in use (running shared-memory bases), its barrier and phase parity
    uint32_t useOff = 0, useBar = barAddr, phases = 0, useBit = 1;
    auto wait_stage = [&]() {
        mbar_wait(useBar, (phases & useBit) ? 1u : 0u);
        phases ^= useBit;
    };
    auto next_stage = [&]() {
        const bool wrap = useOff == (uint32_t)((NST - 1) * STAGE * 8);
        useOff = wrap ? 0u : useOff + (uint32_t)(STAGE * 8);
        useBar = wrap ? barAddr : useBar + 8u;
        useBit = wrap ? 1u : useBit << 1;
    };

    // ---- depth-invariant per-lane constants
    const double zmu = p.zmu[muC], hw = p.hw[muC];
    const double hwG = valid ? hw : 0.0;
    const double bbc0 = cc[p.off_bbc + 2 * laC], bbc1 = cc[p.off_bbc + 2 * laC + 1];
    const double fourPi = 4.0 * kPi;
    const double hw4pi = hwG * fourPi;
    const double r3 = A::rcp(3.0);
    double ca[NSA], cb[NSA], cw[NSA];   // continuum slots: alpha, 2hc/lambda^3, wlamu
#pragma unroll
    for (int tt = 0; tt < NS; ++tt) {
        // The cross-section goes with the record field the lane reads (Vji = g_ij * alpha of wavelength la0 + lsC):
        // an idle lane (ls >= Lw) thus repeats the tile's last wavelength -- a physical opacity -- with zero weights.
        const SlotR &s = T.s[tt];
        const int laD = T.la0 + lsC;
        const int lt = laD - s.Nblue;
        const bool act = laD < Nspect && lt >= 0 && lt < s.Nlam;
        ca[tt] = 0.0;
        cb[tt] = 0.0;
        cw[tt] = 0.0;
        if (!S.kind[tt] && act) {
            ca[tt] = __ldg(p.alpha + s.toff + lt);
            cb[tt] = __ldg(p.twohc + s.toff + lt);
            if (valid) cw[tt] = (__ldg(p.wlacont + s.toff + lt) * hw) * fourPi;
        }
    }
    // segmented sum of J over the mu lanes of a wavelength: lane mu adds the partial sum of lane mu + off when that
    // lane belongs to the same wavelength (a fixed tree)
    bool jadd[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) jadd[q] = (mu + (1 << q)) < NR;

    // Gamma value owned by this lane after the reduce, and where its partial sums live
    const int eOwn = lane / RG;
    const bool writer = (lane % RG) == 0 && eOwn < M && NS > 0;
    double *gp = scr + p.off_part + (size_t)(T.partRow0 + (eOwn < M ? eOwn : 0)) * N + (DIR ? N - 1 : 0);
    const ptrdiff_t jstep = (ptrdiff_t)dk * Nspect;      // the partial-sum pointers advance with the sweep

    SweepT<1, FAST> sw;
    sw.r3 = r3;
    sw.stab = stab;
    double xP = 0.0;   // this lane's J term of the previous step: its mu-sum is taken during the next step

    // second half of the previous step's reductions (its depth: kq): Gamma partial of the lane's matrix entry and the
    // mu-sum of J, rh_method.py:640.  The down and the up sweep write separate partials; the finish kernels add them.
    auto finish_prev = [&]() {
        if constexpr (NS > 0) {
            const double tot = reduce_load<M>(lane, red);
            if (writer) __stcg(gp, tot);
        }
        double sum = xP;
        if constexpr (FAST && NR >= 4 && NR <= 6) {
            // contracted mode: a fixed tree over the wavelength's lanes, (x0 + x1) + (x2 + x3) [+ x4 | + (x4 + x5)] -- three
            // shuffle rounds instead of NR - 1 (no masking: only the leader's sum is used, and its sources are its own
            // wavelength's lanes)
            const double s1 = xP + __shfl_down_sync(0xffffffffu, xP, 1);
            sum = s1 + __shfl_down_sync(0xffffffffu, s1, 2);
            if constexpr (NR == 5) sum += __shfl_down_sync(0xffffffffu, xP, 4);
            if constexpr (NR == 6) sum += __shfl_down_sync(0xffffffffu, s1, 4);
        } else if constexpr (NR <= 6) {   // few angles: pull every other lane's term (no masking; only the leader's sum is used)
#pragma unroll
            for (int m = 1; m < NR; ++m) sum += __shfl_down_sync(0xffffffffu, xP, m);
        } else {                   // many angles: a masked tree
#pragma unroll
            for (int q = 0; (1 << q) < NR; ++q) {
                const double t = __shfl_down_sync(0xffffffffu, sum, 1 << q);
                if (jadd[q]) sum = sum + t;
            }
        }
        if (leader) __stcg(jp, sum);
        gp += dk;
        jp += jstep;
    };

    const int kS = DIR ? N - 1 : 0;
    int k = kS;           // depth of the current step
    // running shared-memory bases of the stage in use: Vij rows / fields / popsT rows
    const double *svL = ring + VOFF + lane, *sfL = ring + FOFF + JW + lsC, *spU = ring + 2 * ST1;
    const double *sv = svL, *sf = sfL, *sp = spU;
#define SET_STAGE()                   \
    do {                              \
        sv = svL + (useOff >> 3);     \
        sf = sfL + (useOff >> 3);     \
        sp = spU + (useOff >> 3);     \
    } while (0)

    // ---- first point (sweep step 0): boundary condition; needs chi at the second point for the upgoing ray
    wait_stage();
    SET_STAGE();
    double chiProbe = 0.0;
    if constexpr (DIR == 1) {   // thermalised lower boundary: chi at kS + dk before the sweep starts (formal_solver.py:205)
        const double *nk = sp + 0 * PW + 1;          // depth N - 2 is the lower of the stage's two popsT rows
        double chiTot = 0.0;
#pragma unroll
        for (int tt = 0; tt < NS; ++tt) {
            const double ni = nk[S.rowI[tt]], nj = nk[S.rowJ[tt]];
            if (S.kind[tt]) {
                const double ld = sv[ST1 + line_index(tt) * kVRow];
                chiTot += A::nmad(nj, T.s[tt].cA * ld, ni * ld);
            } else {
                const double ld = sf[ST1 + (3 + tt) * LW];       // Vji = g_ij * alpha
                chiTot += A::nmad(nj, ld, ni * ca[tt]);
            }
        }
        chiProbe = chiTot + sf[ST1];
    }
    {
#define STEP_KIND 0
#define STEP_SLOT 0
#include "mali_fs_step.inc"
#undef STEP_SLOT
#undef STEP_KIND
    }
    // ---- interior points in pairs: sweep steps 2j + 1 (second half of group j) and 2j + 2 (first half of group j + 1)
    for (int s2 = 2; s2 <= N - 2; s2 += 2) {
        {
#define STEP_KIND 1
#define STEP_SLOT 1
#include "mali_fs_step.inc"
#undef STEP_SLOT
        }
        fetch_group();      // every lane has passed the __syncwarp that ends the step: the stage just left is free
        next_stage();
        wait_stage();
        SET_STAGE();
        {
#define STEP_SLOT 0
#include "mali_fs_step.inc"
#undef STEP_SLOT
#undef STEP_KIND
        }
    }
    // ---- tail: N even: the last point is the second half of the stage in use; N odd: one more interior point
    // there, then the last point opens the final (single-step) group
    if (N & 1) {
        {
#define STEP_KIND 1
#define STEP_SLOT 1
#include "mali_fs_step.inc"
#undef STEP_SLOT
#undef STEP_KIND
        }
        next_stage();
        wait_stage();
        SET_STAGE();
    }
    {
        const int slotLast = (N & 1) ? 0 : 1;
#define STEP_KIND 2
#define STEP_SLOT slotLast
#include "mali_fs_step.inc"
#undef STEP_SLOT
#undef STEP_KIND
    }
    finish_prev();
    if (DIR == 1 && valid) p.I[(size_t)col * p.IStride + (size_t)la * NR + mu] = sw.Iupw;
    if (valid && sw.bad() && p.status != nullptr) atomicOr(p.status + col, 2);
#undef line_index
#undef SET_STAGE
}

// ---- registry of ahead-of-time instances -----------------------------------------------------------------
struct SpecEntry {
    const char *key;
    int nslot;
    int id;
};

}  // namespace mali
