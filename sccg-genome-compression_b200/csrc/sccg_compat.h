// Build switch: nvcc (product, sm_100a) vs. the CPU SIMT emulator used only by the unit tests.
#pragma once

#ifdef SCCG_EMU
// tests/emu/build_emu.sh: g++ -DSCCG_EMU -- kernel LOGIC tests in the GPU-less build container.
#include "simt_emu.h"
#define SCCG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (emu::g_kernel_name = #kernel, emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); }))
#define SCCG_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#define SCCG_SET_MAX_SMEM(kernel, bytes) ((void)0)
// inter-block flags (blocks run on concurrent host threads in the emulator)
#define SCCG_LD_RELAXED_U64(p) __atomic_load_n((const unsigned long long*)(p), __ATOMIC_ACQUIRE)
#define SCCG_ST_RELAXED_U64(p, v) __atomic_store_n((unsigned long long*)(p), (unsigned long long)(v), __ATOMIC_RELEASE)
#else
#include <cuda_runtime.h>
#define SCCG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SCCG_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SCCG_SET_MAX_SMEM(kernel, bytes) \
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
// inter-block flags: value and flag travel in one naturally atomic 64-bit word, so relaxed volatile accesses are enough
#define SCCG_LD_RELAXED_U64(p) (*reinterpret_cast<const volatile unsigned long long*>(p))
#define SCCG_ST_RELAXED_U64(p, v) (*reinterpret_cast<volatile unsigned long long*>(p) = (unsigned long long)(v))
#endif

#include <stdint.h>

#ifdef SCCG_EMU
#define SCCG_NOINLINE __attribute__((noinline))
#else
#define SCCG_NOINLINE __noinline__
#endif

#define SCCG_FULL_MASK 0xffffffffu
