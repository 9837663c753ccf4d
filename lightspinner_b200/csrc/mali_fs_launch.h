// mali_fs_launch.h -- host entry points of the per-class translation units (mali_fs_class.cu), called by mali_api.cu.
#pragma once
#include <cuda_runtime.h>

#include "mali_fs_spec.cuh"

// exact arithmetic (the reference's rounding) and, suffix _fast, contracted arithmetic (mali_device.cuh, Arith)
#define MALI_FS_DECL(SUFFIX)                                                                                          \
    cudaError_t mali_fs_set_attr_##SUFFIX();                                                                          \
    cudaError_t mali_fs_launch_##SUFFIX(const mali::FsCommon &c, const void *tiles /* TileR<2|4|8>[] */, int nt,       \
                                        int ncol, size_t smem, cudaStream_t st, long long *launches);
MALI_FS_DECL(0)
MALI_FS_DECL(1)
MALI_FS_DECL(2)
MALI_FS_DECL(0_fast)
MALI_FS_DECL(1_fast)
MALI_FS_DECL(2_fast)
#undef MALI_FS_DECL
const mali::SpecEntry *mali_fs_registry();
