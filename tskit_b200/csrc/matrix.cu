// matrix.cu -- genotype decode and divergence / relatedness matrices (a9-a12).
#include "plan.cuh"

using namespace tskb;

extern "C" {

int tskb_treeseq_divergence_matrix(const tskb_treeseq_t *self, uint64_t, const uint64_t *,
    const int32_t *, uint64_t, const double *, uint32_t, double *) {
    (void) self;
    return TSKB_ERR_UNSUPPORTED;
}

int tskb_treeseq_genotype_matrix(const tskb_treeseq_t *self, const int32_t *, uint64_t, uint32_t,
    int8_t *) {
    (void) self;
    return TSKB_ERR_UNSUPPORTED;
}

}
