// Dispatcher of the pairwise-intersection back ends.
#include "common.cuh"

using namespace marsb200;

extern "C" int marsb200_pairwise_inter(const uint32_t* bits, int E, int P, int64_t words_per_mask, int32_t* inter,
                                       int backend, void* stream) {
    MARS_REQUIRE(bits && inter, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && P > 0 && words_per_mask > 0 && words_per_mask % 32 == 0, "shape");
    if (backend == MARSB200_PAIR_AUTO)  // tensor cores: block-scaled FP4 when the episode fits one 256-row block, int8 otherwise
        backend = (P <= 256 && words_per_mask * 32 < (1ll << 24)) ? MARSB200_PAIR_FP4 : MARSB200_PAIR_MMA;
    if (backend == MARSB200_PAIR_POPC) return pairwise_popc(bits, E, P, words_per_mask, inter, as_stream(stream));
    if (backend == MARSB200_PAIR_MMA) return pairwise_mma(bits, E, P, words_per_mask, inter, as_stream(stream));
    if (backend == MARSB200_PAIR_FP4) return pairwise_fp4(bits, E, P, words_per_mask, inter, as_stream(stream));
    return fail(MARSB200_ERR_ARG, "%s: unknown backend %lld", "marsb200_pairwise_inter", backend);
}
