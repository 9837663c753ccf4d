// mali_fs_class.cu -- one register class of the structure-specialised formal-solution kernels per translation unit
// (compiled three times, -DMALI_CLS=0|1|2, in parallel: the specialised bodies are most of the library's build time).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "mali_fs_spec.cuh"
#include "mali_fs_launch.h"

#ifndef MALI_CLS
#error "compile with -DMALI_CLS=0|1|2 -DMALI_FAST=0|1"
#endif
#ifndef MALI_FAST
#define MALI_FAST 0
#endif

using namespace mali;

// Ahead-of-time instances of the structure-specialised kernel (generated: tools/gen_spec_instances.py; a
// model-specific library is built with -DMALI_SPEC_INC=\"<file>\" by lightspinner_b200/specialize.py)
#ifndef MALI_SPEC_INC
#define MALI_SPEC_INC "spec_instances.inc"
#endif
#define MALI_SPEC(ID, KEY, ...)                          \
    struct SpecTag##ID {                                 \
        static constexpr TileStruct S = {__VA_ARGS__};   \
    };
#include MALI_SPEC_INC
#undef MALI_SPEC

namespace mali {
// One kernel per register class; the structure id of the tile and the sweep direction select the specialised body
// (a block-uniform switch).  blockIdx.x runs over columns, blockIdx.y over (sweep direction, tile) with the tiles
// sorted by structure: blocks that are resident together share a structure and a direction, hence (nearly always)
// ONE instruction stream per SM -- the depth loop of an instance is 5-12 KB of SASS, and a second or third stream on
// the same SM overflows the 32 KB instruction cache level (ncu: no_instruction was the top stall with interleaved
// directions).  The down and the up sweep of a tile are independent warps (their partial sums go to separate scratch
// copies), which doubles the parallelism of small launches and halves the tail of every wave.
template <int CLS, bool FAST>
__global__ void __launch_bounds__(32, spec_class_warps(CLS)) fs_gamma_kernel_m(const __grid_constant__ MegaParams<CLS> P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // blockIdx.z = slab of colChunk columns, blockIdx.x = column inside the slab, blockIdx.y = (tile, direction): all
    // tiles and both directions of a slab of columns run within a short window, so that what they share -- a column's
    // popsT rows (read by every tile), a tile's fields (read by both directions) -- is found in L2 the second time
    const int nt = gridDim.y >> 1;
    const int dir = P.c.dirInterleave ? (blockIdx.y & 1) : (blockIdx.y >= nt ? 1 : 0);
    const TileR<MegaParams<CLS>::NSP> &T = P.tiles[P.c.dirInterleave ? (blockIdx.y >> 1) : (blockIdx.y - dir * nt)];
    switch (T.spec) {
#define MALI_SPEC(ID, KEY, ...)                                                     \
    case ID:                                                                        \
        if constexpr (spec_class(SpecTag##ID::S.nslot) == CLS) {                    \
            if (dir == 0)                                                           \
                fs_body<SpecTag##ID, MegaParams<CLS>::NSP, 0, FAST>(P.c, T, smem_raw); \
            else                                                                    \
                fs_body<SpecTag##ID, MegaParams<CLS>::NSP, 1, FAST>(P.c, T, smem_raw); \
        }                                                                           \
        break;
#include MALI_SPEC_INC
#undef MALI_SPEC
        default:
            break;
    }
}
}  // namespace mali

#define MALI_CAT2(a, b) a##b
#define MALI_CAT(a, b) MALI_CAT2(a, b)
#if MALI_FAST
#define MALI_FN(stem) MALI_CAT(MALI_CAT(stem, MALI_CLS), _fast)
#else
#define MALI_FN(stem) MALI_CAT(stem, MALI_CLS)
#endif

cudaError_t MALI_FN(mali_fs_set_attr_)()
{
    return cudaFuncSetAttribute(fs_gamma_kernel_m<MALI_CLS, MALI_FAST != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                227 * 1024);
}

cudaError_t MALI_FN(mali_fs_launch_)(const FsCommon &c, const void *tiles, int nt, int ncol, size_t smem, cudaStream_t st,
                                     long long *launches)
{
    using MP = MegaParams<MALI_CLS>;
    using TR = TileR<MP::NSP>;
    static thread_local MP *P = nullptr;   // host-side parameter block (31 KB): reused, every launch copies it
    if (!P) P = new MP();
    P->c = c;
    const TR *src = static_cast<const TR *>(tiles);
    for (int t0 = 0; t0 < nt; t0 += MP::kMaxTiles) {
        const int n = std::min(MP::kMaxTiles, nt - t0);
        memcpy(P->tiles, src + t0, sizeof(TR) * n);
        const int chunk = std::max(1, std::min(c.colChunk, ncol));
        dim3 grid(chunk, 2 * n, (ncol + chunk - 1) / chunk);
        fs_gamma_kernel_m<MALI_CLS, MALI_FAST != 0><<<grid, 32, smem, st>>>(*P);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

#if MALI_CLS == 0 && !MALI_FAST
// registry of the ahead-of-time instances (structure key -> id), used by the host to route tiles
#define MALI_SPEC(ID, KEY, ...) {KEY, SpecTag##ID::S.nslot, ID},
static const SpecEntry kSpecRegistry[] = {
#include MALI_SPEC_INC
    {nullptr, 0, 0}};
#undef MALI_SPEC
const mali::SpecEntry *mali_fs_registry() { return kSpecRegistry; }
#endif
