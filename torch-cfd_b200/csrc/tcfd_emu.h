// Host emulation of the CUDA execution model, for GPU-less debugging of the kernels' index math
// and numerics.  TEST INFRASTRUCTURE ONLY: compiled only into tests/emu/libtcfd_emu.so
// (-DTCFD_EMU, plain g++), never into the product library libtcfd.so and never loaded by the
// torch-cfd_b200 package.  One CTA at a time; its threads are OS threads; __syncthreads() is a
// barrier; "device" memory is host memory.
#pragma once
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define TCFD_HD inline
#define TCFD_D inline

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;
inline unsigned char* emu_smem_ptr = nullptr;
inline unsigned emu_grid_y = 0, emu_grid_z = 0;

struct EmuBarrier {
  std::mutex m;
  std::condition_variable cv;
  unsigned count = 0, waiting = 0, gen = 0;
  void reset(unsigned n) { count = n; waiting = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    unsigned g = gen;
    if (++waiting == count) {
      waiting = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};
inline EmuBarrier emu_barrier;
inline void __syncthreads() { emu_barrier.wait(); }

template <class F>
inline void emu_launch(unsigned grid, unsigned block, size_t smem_bytes, F&& body) {
  std::vector<unsigned char> smem(smem_bytes + 1024 + 64);
  emu_smem_ptr = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 1023) & ~(uintptr_t)1023);
  blockDim.x = block;
  gridDim.x = grid;
  emu_barrier.reset(block);
  std::vector<std::thread> th;
  th.reserve(block);
  for (unsigned t = 0; t < block; ++t) {
    th.emplace_back([&, t] {
      threadIdx.x = t;
      blockIdx.y = emu_grid_y;
      blockIdx.z = emu_grid_z;
      for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        body();
        emu_barrier.wait();  // CTA boundary: all threads finish block b before b+1 reuses smem
      }
    });
  }
  for (auto& x : th) x.join();
  emu_smem_ptr = nullptr;
}

#define TCFD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu_launch((grid), (block), (smem), [&] { kernel(__VA_ARGS__); })
// 3-D grids: the y/z block indices are iterated on the host
#define TCFD_LAUNCH3(kernel, gx, gy, gz, block, smem, stream, ...)          \
  do {                                                                      \
    for (unsigned emu_z = 0; emu_z < (unsigned)(gz); ++emu_z)               \
      for (unsigned emu_y = 0; emu_y < (unsigned)(gy); ++emu_y) {           \
        emu_grid_y = emu_y; emu_grid_z = emu_z;                             \
        gridDim.y = (gy); gridDim.z = (gz);                                 \
        emu_launch((gx), (block), (smem), [&] { kernel(__VA_ARGS__); });    \
      }                                                                     \
    emu_grid_y = emu_grid_z = 0; gridDim.y = gridDim.z = 1;                 \
  } while (0)
#define TCFD_DYN_SMEM(name) unsigned char* name = emu_smem_ptr

// minimal runtime shims
typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
inline cudaError_t cudaMalloc(void** p, size_t n) {  // 256-byte aligned like the device allocator
  *p = std::aligned_alloc(256, ((n ? n : 1) + 255) / 256 * 256);
  return *p ? 0 : 2;
}
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef void* cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* t, cudaEvent_t, cudaEvent_t) { *t = 0.f; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
#define cudaEventDisableTiming 2
#define cudaStreamNonBlocking 1
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
#define cudaHostAllocMapped 2
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return 0; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
