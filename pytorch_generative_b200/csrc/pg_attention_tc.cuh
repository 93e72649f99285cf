// pg_attention_tc.cuh — tensor-core causal attention (placeholder until the tcgen05 kernels land).
#pragma once
namespace {
int attn_fwd_tc(const AttnArgs&, cudaStream_t) {
  pg_set_error("pg_causal_attn_fwd: tcgen05 kernel not built yet");
  return 1;
}
int attn_bwd_tc(const AttnArgs&, cudaStream_t) {
  pg_set_error("pg_causal_attn_bwd: tcgen05 kernel not built yet");
  return 1;
}
}  // namespace
