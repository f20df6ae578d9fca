// host-only check of the chain engines' packet-slice plan (dspi_b200/csrc/chain_streams.cuh): prints "n: bounds..." for a range of
// call lengths; tests/test_slices_cpu.py asserts the invariants.  Built with nvcc (the header pulls in the CUDA runtime API), no GPU used.
#include <cstdio>
#include <cstdint>
#include "chain_streams.cuh"

int main()
{
    for (uint32_t n = 1; n <= 300; n++) {
        uint32_t b[dspi::ChainStreams::kMaxSlices + 1];
        const int k = dspi::ChainStreams::plan_slices(n, b);
        printf("%u:", n);
        for (int i = 0; i <= k; i++) printf(" %u", b[i]);
        printf("\n");
    }
    return 0;
}
