// Probe: does tcgen05.mma have an FP64 kind on sm_100a?  ("to be confirmed with ptxas -arch=sm_100a")
// Build: nvcc -arch=sm_100a -c f64_tcgen05_probe.cu   -- expected to FAIL in ptxas; the message is the result.
#include <cstdint>
__global__ void probe(uint32_t tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n"
                 " tcgen05.mma.cta_group::1.kind::KIND [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc));
}
