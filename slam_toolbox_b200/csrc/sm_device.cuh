// Device-side helpers shared by the scan-matcher kernels.
#pragma once
#include <cstdint>

namespace b200 {

// max-update of one byte with a 32-bit CAS (no native 8-bit atomics). Reads first: almost all
// stamps after the first few land on cells that already hold a larger value.
__device__ __forceinline__ void atomic_max_u8(uint8_t * addr, uint32_t v)
{
  uintptr_t a = reinterpret_cast<uintptr_t>(addr);
  uint32_t * w = reinterpret_cast<uint32_t *>(a & ~uintptr_t(3));
  const int sh = int(a & 3) * 8;
  uint32_t old = *w;
  while (((old >> sh) & 0xFFu) < v) {
    uint32_t nw = (old & ~(0xFFu << sh)) | (v << sh);
    uint32_t prev = atomicCAS(w, old, nw);
    if (prev == old) break;
    old = prev;
  }
}

}  // namespace b200
