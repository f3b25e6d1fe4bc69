// Instantiation unit of the stream-regime tcgen05 GEMM (bfx_tc_gemm.cuh): the three GEMMs of a Dense layer as
// CTA-pair kernels (cluster 2 x 1 x 1, cta_group::2), TMA operand staging, BF16 hi / lo outputs by TMA store.
//
// Reference arithmetic replaced: Flux Dense forward (called at src/likelihoods/feedforward.jl:35,91) and its
// Zygote pullbacks (src/model/BNN.jl:66) for layers wide enough to be real GEMMs (configs[2]: 512 x 512 x 8192).
#include "bfx_tc_gemm.cuh"

#include "bfx_stream.cuh"

namespace bfx {

// A = W (MN-major), B = activations (K-major): act_out = act(W a + b)
cudaError_t tc2_forward(const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B, const tcg::Output& O, cudaStream_t st) {
  return tcg::launch<2, true, false, tcg::EPI_BIAS_ACT>(g, A, B, O, 1, st);
}
// A = W^T (K-major), B = delta (K-major): delta_prev = (W^T delta) .* act'(a)
cudaError_t tc2_dx(const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B, const tcg::Output& O, cudaStream_t st) {
  return tcg::launch<2, false, false, tcg::EPI_MUL_DACT>(g, A, B, O, 1, st);
}
// A = delta (MN-major), B = activations (MN-major), K = batch, split over the batch: dW partials
cudaError_t tc2_dw(const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B, int ksplit, cudaStream_t st) {
  return tcg::launch<2, true, true, tcg::EPI_PLAIN>(g, A, B, tcg::Output{nullptr, nullptr}, ksplit, st);
}

}  // namespace bfx
