// Tile-fused HBM-streaming RK4 for rings too large for one SM's registers (placeholder:
// the kernel lands in a following commit; until then large N runs on qw_generic.cu).
#include "qw_common.cuh"

bool qw_stream_supported(const QwParams &) { return false; }
cudaError_t qw_stream_plan(const QwParams &, qw_plan *) { return cudaErrorNotSupported; }
cudaError_t qw_launch_stream(const QwParams &, cudaStream_t, int64_t *) { return cudaErrorNotSupported; }
