// Probe of the TMEM semantics the fused kernel relies on (run on a B200):
//  * tcgen05.st/ld .32x32b: thread i of warp w addresses TMEM lane 32*(w%4)+i,
//    N consecutive 32-bit columns from the column in taddr;
//  * warps w and w+4 see each other's data in the same lanes after
//    wait::st + fence + CTA barrier + fence.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),
      "=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
    :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),
      "r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]) : "memory");
}
__global__ void k(double* out, double* out2) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_s;
  const uint32_t lane_base = base + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t r[16];
  for (int i = 0; i < 8; ++i) { double d = threadIdx.x + 0.5 * i; r[2*i] = __double2loint(d); r[2*i+1] = __double2hiint(d); }
  tmem_st16(lane_base + 400 + 16 * (warp >> 2), r);     // high columns too: [400,432)
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t q[16];
  tmem_ld16(lane_base + 400 + 16 * (warp >> 2), q);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = __hiloint2double(q[2*i+1], q[2*i]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  tmem_ld16(lane_base + 400 + 16 * ((warp >> 2) ^ 1), q);  // what the partner warp (w ^ 4) wrote
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out2[threadIdx.x * 8 + i] = __hiloint2double(q[2*i+1], q[2*i]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "n"(512));
  if (threadIdx.x == 0) out[2048] = (double)base;
}
int main() {
  double *d, *d2; cudaMalloc(&d, 2049 * 8); cudaMalloc(&d2, 2048 * 8);
  k<<<1, 256>>>(d, d2);
  static double h[2049], h2[2048];
  cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2, d2, sizeof h2, cudaMemcpyDeviceToHost);
  int bad = 0, bad2 = 0;
  for (int t = 0; t < 256; ++t)
    for (int i = 0; i < 8; ++i) {
      if (h[t * 8 + i] != t + 0.5 * i) ++bad;
      if (h2[t * 8 + i] != (t ^ 128) + 0.5 * i) ++bad2;
    }
  printf("tmem_probe: cuda=%s base=%g own_bad=%d partner_bad=%d\n", cudaGetErrorString(e), h[2048], bad, bad2);
  return bad + bad2;
}
