// TMA (cp.async.bulk.tensor) + mbarrier wrappers used by the cols kernel to pull a strided
// [kx][TC columns] tile of H into shared memory asynchronously (SASS: UTMALDG + SYNCS), with the
// hardware 128-byte swizzle so that the column reads that follow are bank-conflict free.
// Under TCFD_EMU the same interface is a synchronous strided copy that applies the same swizzle.
#pragma once
#include "tcfd_common.cuh"

#ifndef TCFD_EMU
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif

namespace tcfd {

// The H tile source.  H holds, per sample, the N x N array z[kx][y] of packed complex entries (ENT bytes):
// lanes (u + i v, dw/dx + i dw/dy) after the y-inverse -- i.e. exactly the input of the packed x-axis
// inverse transform of column y (rows kx > N/2 are written by the rows units as conjugates, so the column
// kernel reads every entry once and does no Hermitian unpacking).  A tile is, for every row kx of one
// sample, the 4 entries of 4 consecutive columns y.
// The advection rows of the dataflow schedule (ns2d_flow.cuh) live in blocks of 8 rows, [slot][row / 8][y][row % 8]
// entries: a cols item writes, for each of its 4 columns y, the entries of ALL rows -- with 8 consecutive rows adjacent
// a warp's store covers 4 cache lines instead of 29 (ncu: these stores were 43 % of the kernel's global-memory tag
// requests) -- and a rows unit pulls its row (N entries, one per 8-entry line) with a rank-4 tiled copy (`adv`).
#ifndef TCFD_EMU
struct alignas(64) TileMaps {
  CUtensorMap main;  // rank 3, box {4 entries, min(256, N) rows, 1 sample}
  CUtensorMap adv;   // rank 4 {4 reals, 8 rows of a block, N columns y, slots x blocks}, box {4, 1, min(256, N), 1}
};
#else
struct TileMaps {
  const unsigned char* base;
  size_t row_bytes, sample_bytes;
  const unsigned char* adv_base;
};
#endif
constexpr int ADV_BLOCK = 8;

// Shared-memory image of a tile with NR rows whose inner box is IB = 64 or 128 bytes (hardware swizzle of
// the same width).  The rows arrive in boxes of BOX rows, stored [box][row][IB] (i.e. row-major).
// chunk_offset returns the byte offset of 16-byte chunk j of a row -- the swizzle XORs the chunk index with
// address bits 7.. (the tile base is 1 KB aligned, so offsets and addresses agree in those bits).
template <int NR, int IB>
struct TileGeom {
  static constexpr int BOX = NR < 256 ? NR : 256;
  static constexpr int NBOX = NR / BOX;
  static constexpr int BYTES = ((NR * IB) + 1023) / 1024 * 1024;
  static constexpr int XMASK = IB / 16 - 1;  // 3 (64B swizzle) or 7 (128B swizzle)
  TCFD_HD static int chunk_offset(int row, int j) {
    const int ro = row * IB;
    return ro + ((j ^ ((ro >> 7) & XMASK)) << 4);
  }
};

// offset of a shared-memory pointer inside the CTA's shared window
TCFD_D unsigned smem_offset(const void* p) {
#ifndef TCFD_EMU
  return (unsigned)__cvta_generic_to_shared(p);
#else
  return (unsigned)(reinterpret_cast<uintptr_t>(p) & 0xffffffffu);
#endif
}

#ifndef TCFD_EMU
TCFD_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

TCFD_D void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
TCFD_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Wait for the phase.  try_wait suspends the warp for the hardware's default window; the retry loop is four
// instructions and bounded by the retry count, so that a lost transaction traps instead of hanging the GPU (an earlier
// version re-read the SM clock on every retry, which cost issue slots while other warps had work; an explicit 2 us
// suspend-time hint doubled the time of the spectral-conv plane kernels at Y = 128: their waits are short and frequent).
TCFD_D bool mbar_try_wait(unsigned addr, unsigned phase) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(phase)
      : "memory");
  return done != 0;
}
TCFD_D void mbar_wait(unsigned long long* bar, unsigned phase) {
  const unsigned addr = smem_u32(bar);
  if (mbar_try_wait(addr, phase)) return;
  for (unsigned spin = 0; !mbar_try_wait(addr, phase); ++spin)
    if (spin > (1u << 26)) asm volatile("trap;");
}
TCFD_D void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2),
      "r"(smem_u32(bar))
      : "memory");
}
TCFD_D void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(smem_u32(bar))
      : "memory");
}
TCFD_D void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(map)) : "memory");
}
#endif

// 1-D bulk copy global -> shared (SASS: UBLKCP); src, dst and bytes multiples of 16.  The caller
// announces the total byte count of all copies of a stage with stage_expect().
TCFD_D void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
#ifndef TCFD_EMU
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
#else
  (void)bar;
  std::memcpy(dst, src, bytes);
#endif
}
// Per-thread asynchronous 16-byte copy global -> shared (SASS: LDGSTS, L2 only): unlike the bulk copies every lane has
// its own source and destination, so a scattered destination layout (padded rows) costs one instruction per 16 bytes
// of a thread instead of one warp-serialised UBLKCP per piece.  Completion: ldgsts_wait_all() by the issuing thread,
// then a barrier with the readers.
TCFD_D void ldgsts16(void* dst, const void* src) {
#ifndef TCFD_EMU
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(src)) : "memory");
#else
  std::memcpy(dst, src, 16);
#endif
}
TCFD_D void ldgsts_commit() {
#ifndef TCFD_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
TCFD_D void ldgsts_wait_all() {
#ifndef TCFD_EMU
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// 1-D bulk copy shared -> global (SASS: UBLKCP ... bulk_group), committed as one group.  The writers
// of the shared source must have executed fence_async_smem() and synchronised with the issuing thread.
TCFD_D void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
#ifndef TCFD_EMU
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<unsigned long long>(gdst)), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#else
  std::memcpy(gdst, ssrc, bytes);
#endif
}
// wait until at most PENDING of this thread's bulk-store groups are still READING their source
template <int PENDING>
TCFD_D void bulk_store_wait_read() {
#ifndef TCFD_EMU
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
#endif
}
TCFD_D void fence_async_smem() {  // generic-proxy writes to shared memory -> visible to the async proxy
#ifndef TCFD_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
TCFD_D void stage_expect(unsigned long long* bar, unsigned bytes) {
#ifndef TCFD_EMU
  mbar_expect_tx(bar, bytes);
#else
  (void)bar;
  (void)bytes;
#endif
}
TCFD_D void stage_barrier_init(unsigned long long* bar) {
#ifndef TCFD_EMU
  mbar_init(bar, 1);
#else
  (void)bar;
#endif
}

// Issue the load of one tile (called by ONE thread): columns y0..y0+3 of `sample`, all NR rows.
template <int NR, int IB>
TCFD_D void tile_load_issue(unsigned char* tile, const TileMaps& maps, int y0, int sample, unsigned long long* bar) {
  typedef TileGeom<NR, IB> G;
#ifndef TCFD_EMU
  mbar_expect_tx(bar, (unsigned)NR * (unsigned)IB);
  const int c0 = y0 * 4;  // innermost coordinate in reals: 4 reals per entry
#pragma unroll
  for (int b = 0; b < G::NBOX; ++b) tma_load_3d(tile + b * G::BOX * IB, &maps.main, c0, b * G::BOX, sample, bar);
#else
  (void)bar;
  const int ent = IB / 4;  // bytes per entry
  for (int r = 0; r < NR; ++r) {
    const unsigned char* src = maps.base + (size_t)sample * maps.sample_bytes + (size_t)r * maps.row_bytes + (size_t)y0 * ent;
    for (int j = 0; j < IB / 16; ++j) std::memcpy(tile + G::chunk_offset(r, j), src + 16 * j, 16);
  }
#endif
}

// Issue the load of advection row `row` of block-row index `blk` (= slot * blocks_per_slot + row / 8): NR entries of ENT
// bytes into dst (called by ONE thread, after stage_expect on the same barrier).
template <int NR, int ENT>
TCFD_D void adv_row_issue(unsigned char* dst, const TileMaps& maps, int row_in_block, int blk, unsigned long long* bar) {
  constexpr int BOX = NR < 256 ? NR : 256;
#ifndef TCFD_EMU
#pragma unroll
  for (int b = 0; b < NR / BOX; ++b) tma_load_4d(dst + (size_t)b * BOX * ENT, &maps.adv, 0, row_in_block, b * BOX, blk, bar);
#else
  (void)bar;
  const unsigned char* src = maps.adv_base + ((size_t)blk * NR * ADV_BLOCK + row_in_block) * ENT;
  for (int y = 0; y < NR; ++y) std::memcpy(dst + (size_t)y * ENT, src + (size_t)y * ADV_BLOCK * ENT, ENT);
#endif
}

// All threads of the CTA wait for the tile whose load was issued with `bar` (phase = parity of
// the number of tiles consumed so far on this barrier).
TCFD_D void tile_load_wait(unsigned long long* bar, unsigned phase) {
#ifndef TCFD_EMU
  mbar_wait(bar, phase);
#else
  (void)bar;
  (void)phase;
  __syncthreads();  // the emulated load is a plain copy by one host thread
#endif
}

}  // namespace tcfd
