// mali_fs_launch.h -- host entry points of the per-class translation units (mali_fs_class.cu), called by mali_api.cu.
#pragma once
#include <cuda_runtime.h>

#include "mali_fs_spec.cuh"

cudaError_t mali_fs_set_attr_0();
cudaError_t mali_fs_set_attr_1();
cudaError_t mali_fs_set_attr_2();
// tiles: array of TileR<2> / TileR<4> / TileR<8>
cudaError_t mali_fs_launch_0(const mali::FsCommon &c, const void *tiles, int nt, int ncol, size_t smem, cudaStream_t st,
                             long long *launches);
cudaError_t mali_fs_launch_1(const mali::FsCommon &c, const void *tiles, int nt, int ncol, size_t smem, cudaStream_t st,
                             long long *launches);
cudaError_t mali_fs_launch_2(const mali::FsCommon &c, const void *tiles, int nt, int ncol, size_t smem, cudaStream_t st,
                             long long *launches);
const mali::SpecEntry *mali_fs_registry();
