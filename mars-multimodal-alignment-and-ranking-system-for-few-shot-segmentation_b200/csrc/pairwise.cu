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

// One pixel slice of the intersections: the single-episode schedule packs the masks slice by slice and counts the
// intersections of slice k on the tensor cores while slice k + 1 is still streaming in from HBM.
extern "C" int marsb200_pairwise_inter_slice(const uint32_t* bits, int E, int P, int64_t words_per_mask, int64_t word_begin,
                                             int64_t word_count, int accumulate, int32_t* inter, int backend, void* stream) {
    MARS_REQUIRE(bits && inter, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && P > 0 && words_per_mask > 0 && words_per_mask % 32 == 0, "shape");
    MARS_REQUIRE(word_begin >= 0 && word_count > 0 && word_begin + word_count <= words_per_mask, "slice outside the mask");
    MARS_REQUIRE(word_begin % 8 == 0 && word_count % 8 == 0, "slice must be whole 256-pixel blocks");
    if (backend == MARSB200_PAIR_AUTO)
        backend = (P <= 256 && words_per_mask * 32 < (1ll << 24)) ? MARSB200_PAIR_FP4 : MARSB200_PAIR_MMA;
    if (backend == MARSB200_PAIR_MMA)
        return pairwise_mma(bits, E, P, words_per_mask, inter, as_stream(stream), word_begin, word_count, accumulate != 0);
    if (backend == MARSB200_PAIR_FP4)
        return pairwise_fp4(bits, E, P, words_per_mask, inter, as_stream(stream), word_begin, word_count, accumulate != 0);
    return fail(MARSB200_ERR_UNSUPPORTED, "%s: pixel slices exist for the tensor-core back ends only (backend %lld)",
                "marsb200_pairwise_inter_slice", backend);
}
