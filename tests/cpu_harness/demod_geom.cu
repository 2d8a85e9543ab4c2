// Host-only harness for tests/test_search_params.py: compile-time geometry of demod_fast_kernel for every
// specialised chip length (rtlamr_b200/csrc/demod_fast.cuh).  No kernel is launched.
#include <cstdio>

#include "../../rtlamr_b200/csrc/demod_fast.cuh"

template <int CL>
void row(bool first) {
    using G = ert::FastGeom<CL>;
    using G3 = ert::FastGeom<CL, 3>;
    const int W = ert::fast_warps<CL>();
    printf("%s{\"CL\":%d,\"L\":%d,\"pad\":%d,\"row_bytes\":%d,\"stages\":%d,\"warps\":%d,\"packed\":%s,\"tail_bits\":%d,"
           "\"smem\":%d,\"smem7\":%d,\"smem3\":%d,\"variant\":%d}",
           first ? "" : ",", CL, G::L, G::kPad, G::kRowBytes, G::kStages, W, G::kPacked ? "true" : "false", G::kTailBits,
           ert::fast_smem_bytes<G>(W, 1024), ert::fast_smem_bytes<G>(W == 8 ? 7 : W, 1024), ert::fast_smem_bytes<G3>(W, 1024),
           ert::demod_fast_variant(CL, 4096));
}

int main() {
    printf("[");
    row<32>(true); row<40>(false); row<48>(false); row<56>(false); row<64>(false);
    row<72>(false); row<78>(false); row<80>(false); row<88>(false); row<96>(false);
    printf("]\n");
    return 0;
}
