// placeholder: replaced by the real kernels
#include "graph_handle.hpp"
extern "C" {
int poasta_align_batch(int, const poasta_graph_view *, uint32_t, const uint8_t *const *, const uint32_t *, const uint32_t *,
                       uint32_t, const poasta_params *, poasta_result *) { return POASTA_ERR_CUDA; }
int poasta_last_batch_stats(double *, uint32_t *, uint64_t *) { return POASTA_ERR_CUDA; }
int poasta_msa_batch(int, const char *const *, const size_t *, uint32_t, const poasta_params *, poasta_graph **, uint32_t *,
                     uint64_t *) { return POASTA_ERR_CUDA; }
int poasta_msa_batch_into(int, poasta_graph **, const char *const *, const size_t *, uint32_t, const poasta_params *, uint32_t *,
                          uint64_t *) { return POASTA_ERR_CUDA; }
}
