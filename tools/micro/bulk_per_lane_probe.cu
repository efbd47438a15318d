#include <cstdint>
__global__ void k(const unsigned long long* __restrict__ g, const unsigned* __restrict__ start, unsigned long long* out) {
  __shared__ __align__(128) unsigned char sm[32 * 128];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x;
  unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(bar_a));
  __syncwarp();
  const unsigned long long* src = g + start[lane];
  unsigned dst = (unsigned)__cvta_generic_to_shared(sm + lane * 128);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 128;" ::"r"(bar_a) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
               ::"r"(dst), "l"(src), "r"(bar_a) : "memory");
  unsigned ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a) : "memory");
  out[lane] = reinterpret_cast<unsigned long long*>(sm + lane * 128)[3];
}
