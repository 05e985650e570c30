// tcgen05 / TMEM / TMA arm of the conv engine (NIC_PREC_BF16, NIC_PREC_BF16X3) - under construction.
#include "conv_common.cuh"

namespace nic {

int conv_fwd_tc(const nic_conv_desc*, const void*, const void*, const float*, const void*, const float*, void*, void*, size_t, cudaStream_t) {
  return fail(NIC_E_UNSUPPORTED, "conv: tcgen05 arm not built yet");
}
int pack_weight_tc(const nic_conv_desc*, const TapTable&, const float*, void*, cudaStream_t) {
  return fail(NIC_E_UNSUPPORTED, "pack_conv_weight: tcgen05 arm not built yet");
}
int pack_gdn_tc(int32_t, float, const float*, const float*, float*, void*, int32_t, cudaStream_t) {
  return fail(NIC_E_UNSUPPORTED, "pack_gdn: tcgen05 arm not built yet");
}
size_t packed_weight_elems_tc(const nic_conv_desc*, const TapTable&) { return 0; }
size_t conv_workspace_bytes_tc(const nic_conv_desc*) { return 0; }

}  // namespace nic
