// lg_compat.h - the one place where the kernel sources see either real CUDA (nvcc, sm_100a: the
// product build) or the test-only SIMT emulator (g++ -DLG_EMULATE: tests/emu, CPU development box).
#pragma once
#ifdef LG_EMULATE
#include "simt_emu.h"
#define LG_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::g_blk->smem)
#define LG_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(emu::dim3_t(grid), emu::dim3_t(block), (smem), [=]() { kern(__VA_ARGS__); })
#define LG_HD
#define LG_NAMED_BARRIER(id, count) emu::named_barrier((id), (count))
#else
#include <cuda_runtime.h>
#define LG_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char lg_smem_raw[];  \
    type *name = reinterpret_cast<type *>(lg_smem_raw)
#define LG_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LG_HD __host__ __device__
#define LG_NAMED_BARRIER(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#endif
#define LG_FULL 0xffffffffu
