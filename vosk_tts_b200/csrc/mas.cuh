// mas.cuh -- Monotonic Alignment Search on the GPU (the reference runs it on the host: the score matrix is copied to the
// CPU, a Cython loop per batch item under prange, and the path is copied back -- training/vits2/monotonic_align/
// __init__.py:6-22, core.pyx:7-43).  Same recurrence, same evaluation order per cell, fp32 adds only => bit-identical paths.
//
//   forward (core.pyx:16-29):  value[y, x] += max(value[y-1, x-1], value[y-1, x])   for max(0, t_x + y - t_y) <= x < min(t_x, y + 1)
//                              with value[-1, -1] := 0, everything else outside the band := -1e9
//   backtrack (core.pyx:31-34): index = t_x - 1; for y = t_y-1 .. 0: path[y, index] = 1;
//                              if index != 0 and (index == y or value[y-1, index] < value[y-1, index-1]): index -= 1
//
// One CTA per batch item; a row of the band is computed by all threads in parallel (each cell depends on two cells of the
// previous row only), the previous row is kept in shared memory (ping-pong), the raw scores of the next row are prefetched
// into registers before the row barrier.  HBM-bound integer/float work: T_y * T_x * 4 bytes read + written once, plus the
// path (zero-filled by the caller's memset, T_y ones written).  The backtrack is a dependent chain of T_y steps of one thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vtts {

constexpr int MAS_THREADS = 256;
constexpr int MAS_MAXPT = 8;        // columns per thread: T_x <= MAS_THREADS * MAS_MAXPT

__global__ void __launch_bounds__(MAS_THREADS)
mas_kernel(float* __restrict__ value, int* __restrict__ path, const int* __restrict__ t_ys, const int* __restrict__ t_xs, int Ty, int Tx) {
  extern __shared__ float mas_rows[];            // [2][Tx]
  const int b = blockIdx.x;
  float* v = value + (size_t)b * Ty * Tx;
  int* p = path + (size_t)b * Ty * Tx;
  const int t_y = min(t_ys[b], Ty), t_x = min(t_xs[b], Tx);
  if (t_y <= 0 || t_x <= 0) return;
  const float MAXNEG = -1e9f;
  const int tid = threadIdx.x;
  float raw[MAS_MAXPT];                           // scores of the row about to be processed, columns tid + k * MAS_THREADS
#pragma unroll
  for (int k = 0; k < MAS_MAXPT; ++k) {
    const int x = tid + k * MAS_THREADS;
    raw[k] = x < t_x ? v[x] : 0.f;
  }
  for (int y = 0; y < t_y; ++y) {
    float* cur = mas_rows + (y & 1) * Tx;
    const float* prv = mas_rows + ((y & 1) ^ 1) * Tx;
    const int lo = max(0, t_x + y - t_y), hi = min(t_x, y + 1);
    float nxt[MAS_MAXPT];
#pragma unroll
    for (int k = 0; k < MAS_MAXPT; ++k) {         // prefetch row y + 1 (independent of this row's results)
      const int x = tid + k * MAS_THREADS;
      nxt[k] = (x < t_x && y + 1 < t_y) ? v[(size_t)(y + 1) * Tx + x] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < MAS_MAXPT; ++k) {
      const int x = tid + k * MAS_THREADS;
      if (x >= lo && x < hi) {
        const float v_cur = (x == y) ? MAXNEG : prv[x];
        const float v_prev = (x == 0) ? (y == 0 ? 0.f : MAXNEG) : prv[x - 1];
        const float nv = __fadd_rn(raw[k], v_prev > v_cur ? v_prev : v_cur);
        v[(size_t)y * Tx + x] = nv;
        cur[x] = nv;
      }
    }
#pragma unroll
    for (int k = 0; k < MAS_MAXPT; ++k) raw[k] = nxt[k];
    __syncthreads();                              // row y complete (shared + this CTA's global writes) before row y + 1 reads it
  }
  if (tid == 0) {
    int index = t_x - 1;
    for (int y = t_y - 1; y >= 0; --y) {
      p[(size_t)y * Tx + index] = 1;
      if (index != 0 && (index == y || v[(size_t)(y - 1) * Tx + index] < v[(size_t)(y - 1) * Tx + index - 1])) --index;
    }
  }
}

}  // namespace vtts
