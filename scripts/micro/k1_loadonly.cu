// What does the memory system deliver for K1's access pattern when the arithmetic is taken away?
// One warp per tile of 4096 shots of one PUB (obs rows 8 B/shot, qpd 1 B/shot), tiles strided over the grid exactly as
// in qcut_tally_sliced; every loaded word is XOR-ed into one register so that the loads cannot be dropped.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o k1_loadonly k1_loadonly.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct u4 { uint32_t x, y, z, w; };
template <int MODE>
__device__ __forceinline__ u4 ld16(const uint8_t* p) {
  u4 v;
  if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// PUB p: obs at p * pub_stride, qpd at p * pub_stride + shots * 8 (16-byte aligned), like the staged layout
// QPD: 0 = no qpd loads, 1 = 16 x LDG.U16 per step (as K1), 2 = one LDG.128 per step and half the lanes (same bytes)
// STEPS_IN_FLIGHT: 1 = 16 loads then consume (as K1); 2 = two steps of loads back to back before consuming
template <int MODE, int QPD, int INFLIGHT, int MINB>
__global__ void __launch_bounds__(256, MINB)
loadonly(const uint8_t* __restrict__ data, long long n_tiles, int tiles_per_pub, long long pub_stride, int shots,
         unsigned* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  uint32_t acc = 0;
  for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles; tile += n_warps) {
    const long long pub = tile / tiles_per_pub;
    const int tip = (int)(tile - pub * tiles_per_pub);
    const uint8_t* obs = data + pub * pub_stride + (long long)tip * 4096 * 8;
    const uint8_t* qpd = data + pub * pub_stride + (long long)shots * 8 + (long long)tip * 4096;
#pragma unroll 1
    for (int s = 0; s < 4; s += INFLIGHT) {
      u4 v[16 * INFLIGHT];
#pragma unroll
      for (int i = 0; i < INFLIGHT; ++i)
#pragma unroll
        for (int k = 0; k < 16; ++k) v[16 * i + k] = ld16<MODE>(obs + (s + i) * 8192 + lane * 16 + k * 512);
      uint32_t q[16 * INFLIGHT];
      if (QPD == 1) {
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i)
#pragma unroll
          for (int k = 0; k < 16; ++k) q[16 * i + k] = __ldg(reinterpret_cast<const unsigned short*>(qpd + (s + i) * 1024 + lane * 2 + k * 64));
      } else if (QPD == 2) {
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i) {
          u4 t = ld16<MODE>(qpd + (s + i) * 1024 + (lane & 15) * 16 + (lane >> 4) * 512);
          q[16 * i] = t.x ^ t.y ^ t.z ^ t.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 16 * INFLIGHT; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
      if (QPD == 1) {
#pragma unroll
        for (int k = 0; k < 16 * INFLIGHT; ++k) acc ^= q[k];
      } else if (QPD == 2) {
#pragma unroll
        for (int i = 0; i < INFLIGHT; ++i) acc ^= q[16 * i];
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;  // never true in practice; keeps the loads alive
}

template <int MODE, int QPD, int INFLIGHT, int MINB>
void run(const char* name, const uint8_t* d, long long n_pubs, int shots, long long pub_stride, unsigned* out, int ctas_per_sm, int smem_pad) {
  const int tiles_per_pub = shots / 4096;
  const long long n_tiles = n_pubs * tiles_per_pub;
  auto k = loadonly<MODE, QPD, INFLIGHT, MINB>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pad));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = 148 * ctas_per_sm;
  for (int i = 0; i < 3; ++i) k<<<grid, 256, smem_pad>>>(d, n_tiles, tiles_per_pub, pub_stride, shots, out);
  CK(cudaDeviceSynchronize());
  const int reps = 10;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k<<<grid, 256, smem_pad>>>(d, n_tiles, tiles_per_pub, pub_stride, shots, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  const double bytes = (double)n_pubs * shots * (8.0 + (QPD ? 1.0 : 0.0));
  printf("%-44s ctas/sm %d  %.3f ms  %.0f GB/s  (%.3f of 6551)\n", name, ctas_per_sm, ms, bytes / ms * 1e-6, bytes / ms * 1e-6 / 6551.0);
  fflush(stdout);
}

int main() {
  const long long n_pubs = 97000;  // one subsystem of cfg4: 97 000 PUBs x 8192 shots
  const int shots = 8192;
  const long long pub_stride = (long long)shots * 9;  // obs (64 KB) then qpd (8 KB)
  const size_t total = (size_t)n_pubs * pub_stride + 4096;
  uint8_t* d = nullptr; unsigned* out = nullptr;
  CK(cudaMalloc(&d, total)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(d, 0x5a, total));
  // smem_pad limits residency the way 128 registers do: 2 CTAs/SM need no pad with MINB = 2 and grid = 296
  run<0, 1, 1, 2>("as K1 (nc no_allocate, qpd 16xU16)", d, n_pubs, shots, pub_stride, out, 2, 0);
  run<0, 0, 1, 2>("no qpd loads", d, n_pubs, shots, pub_stride, out, 2, 0);
  run<0, 2, 1, 2>("qpd as one LDG.128 per step", d, n_pubs, shots, pub_stride, out, 2, 0);
  run<1, 1, 1, 2>("nc, L1 allocating", d, n_pubs, shots, pub_stride, out, 2, 0);
  run<2, 1, 1, 2>("ld.global.cs", d, n_pubs, shots, pub_stride, out, 2, 0);
  run<0, 1, 2, 1>("two steps in flight, 1 CTA/SM", d, n_pubs, shots, pub_stride, out, 1, 0);
  run<0, 1, 2, 2>("two steps in flight, 2 CTAs/SM", d, n_pubs, shots, pub_stride, out, 2, 0);
  run<0, 1, 1, 3>("as K1, 3 CTAs/SM", d, n_pubs, shots, pub_stride, out, 3, 0);
  run<0, 1, 1, 4>("as K1, 4 CTAs/SM", d, n_pubs, shots, pub_stride, out, 4, 0);
  run<0, 1, 1, 8>("as K1, 8 CTAs/SM", d, n_pubs, shots, pub_stride, out, 8, 0);
  run<0, 0, 1, 8>("no qpd, 8 CTAs/SM", d, n_pubs, shots, pub_stride, out, 8, 0);
  run<0, 1, 1, 1>("as K1, 1 CTA/SM", d, n_pubs, shots, pub_stride, out, 1, 0);
  return 0;
}
